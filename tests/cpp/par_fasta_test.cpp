// CPU check of host/par_fasta.hpp: chunked, multi-threaded parsing hands out exactly the records of a serial parse, in
// file order, for FASTA texts with every reader quirk (multi-line records, blank lines, blanks inside lines, lower case,
// '>' inside a sequence line, text before the first header, no trailing newline, empty records, tiny chunk sizes).
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include "par_fasta.hpp"
#include "seq_file.hpp"

using namespace tgio;

static void parse_all(const char* d, size_t n, RecordBatch& rb) {
    InchwormFastaReader rd(d, n);
    const char* h; size_t hl;
    while (true) {
        const size_t before = rb.recs.size();
        if (!rd.next(&h, &hl, rb.recs)) break;
        if (rb.recs.size() == before) continue;
        const char* acc; size_t al;
        accession_of(h, hl, &acc, &al);
        rb.end_record();
        rb.add_name(acc, al);
    }
}

int main() {
    srand(7);
    long checked = 0;
    for (int trial = 0; trial < 300; trial++) {
        std::string t;
        if (trial % 5 == 0) t += "junk before the first header\nmore junk\n";
        const int nrec = rand() % 40;
        for (int r = 0; r < nrec; r++) {
            t += ">r" + std::to_string(r) + (rand() % 3 ? " desc text" : "") + "\n";
            const int lines = rand() % 4;
            for (int l = 0; l < lines; l++) {
                const int len = rand() % 70;
                for (int i = 0; i < len; i++) {
                    const int x = rand() % 40;
                    t += x == 0 ? ' ' : x == 1 ? '\t' : x == 2 ? '>' : x == 3 ? 'n' : "ACGTacgt"[rand() & 7];
                }
                // a '>' may appear INSIDE a line, never at its start here (that would be a header by definition)
                if (!t.empty() && t.back() == '\n') {}
                t += "\n";
                if (rand() % 9 == 0) t += "\n";
            }
        }
        // make sure no sequence line starts with '>' (our generator could have produced one): prefix with 'A'
        for (size_t i = 1; i + 1 < t.size(); i++)
            if (t[i] == '>' && t[i - 1] == '\n' && !(t[i + 1] == 'r')) t[i] = 'A';
        if (trial % 4 == 0 && !t.empty() && t.back() == '\n') t.pop_back();      // no trailing newline
        {   // the same for the reader `jellyfish count` uses (seq_file.hpp: lines of a record joined, '\r' dropped)
            std::string tj = t;
            for (size_t i = 0; i + 1 < tj.size(); i++) if (tj[i + 1] == '\n' && rand() % 6 == 0 && tj[i] != '\n' && tj[i] != '>') tj[i] = '\r';
            std::vector<char> one;
            parse_sequence_file(tj.data(), tj.data() + tj.size(), one, (size_t)-1, [] {});
            for (size_t target : {size_t(1), size_t(23), size_t(100000)}) {
                OrderedChunkParser p(tj.data(), tj.size(), target, 4, 6, [](const char* d, size_t n, RecordBatch& rb) {
                    parse_sequence_file(d, d + n, rb.recs, (size_t)-1, [] {});
                });
                std::vector<char> cat;
                RecordBatch rb;
                while (p.next(rb)) cat.insert(cat.end(), rb.recs.begin(), rb.recs.end());
                if (cat != one) { fprintf(stderr, "MISMATCH (sequence-file reader) trial %d target %zu\n", trial, target); return 1; }
            }
        }
        {   // the Chrysalis reader (ReadsToTranscripts): the line after a header is sequence WHATEVER it starts with, so the
            // text gets '>' lines in runs -- header, '>'-sequence, header, ... -- and the chunks may only be cut at a '>'
            // line that follows a plain line (fasta_chunks(.., after_sequence_line = true))
            std::string tc;
            const int nrec2 = rand() % 30;
            for (int r = 0; r < nrec2; r++) {
                tc += ">c" + std::to_string(r) + (rand() % 2 ? " x y" : "") + "\n";
                const int kind = rand() % 6;
                if (kind == 0) tc += ">looks like a header but is the sequence\n";
                else if (kind == 1) { tc += ">again\n"; tc += "ACGT\n"; }
                else if (kind == 2) tc += "\n";                                           // empty first sequence line
                else { const int lines = 1 + rand() % 3; for (int l = 0; l < lines; l++) { const int len = rand() % 60; for (int i = 0; i < len; i++) tc += "ACGTNacgt"[rand() % 9]; tc += "\n"; } }
            }
            if (trial % 3 == 0 && !tc.empty()) tc.pop_back();                               // unterminated last line: lost / record dropped
            auto parse_dna = [](const char* d, size_t n, RecordBatch& rb) {
                DnaStreamReader rd(d, n);
                const char* name; size_t name_len;
                while (rd.next(&name, &name_len, rb.recs)) { rb.end_record(); rb.add_name(name, name_len); }
            };
            RecordBatch one;
            parse_dna(tc.data(), tc.size(), one);
            for (size_t target : {size_t(1), size_t(9), size_t(64), size_t(100000)}) {
                OrderedChunkParser p(tc.data(), tc.size(), target, 4, 6, parse_dna, /*after_sequence_line=*/true);
                RecordBatch all, rb;
                while (p.next(rb)) {
                    for (size_t i = 0; i < rb.count(); i++) {
                        all.recs.insert(all.recs.end(), rb.seq(i), rb.seq(i) + rb.seq_len(i));
                        all.end_record();
                        all.add_name(rb.name(i), rb.name_len(i));
                    }
                }
                if (all.recs != one.recs || all.offs != one.offs || all.names != one.names || all.name_offs != one.name_offs) {
                    fprintf(stderr, "MISMATCH (Chrysalis reader) trial %d target %zu\n", trial, target);
                    return 1;
                }
            }
        }
        RecordBatch serial;
        parse_all(t.data(), t.size(), serial);
        for (size_t target : {size_t(1), size_t(17), size_t(200), size_t(100000)}) {
            for (unsigned threads : {1u, 3u, 8u}) {
                OrderedChunkParser p(t.data(), t.size(), target, threads, 3, parse_all);
                RecordBatch all, rb;
                while (p.next(rb)) {
                    for (size_t i = 0; i < rb.count(); i++) {
                        all.recs.insert(all.recs.end(), rb.seq(i), rb.seq(i) + rb.seq_len(i));
                        all.end_record();
                        all.add_name(rb.name(i), rb.name_len(i));
                    }
                }
                if (all.recs != serial.recs || all.offs != serial.offs || all.names != serial.names || all.name_offs != serial.name_offs) {
                    fprintf(stderr, "MISMATCH trial %d target %zu threads %u\n", trial, target, threads);
                    return 1;
                }
                checked += (long)all.count();
            }
        }
    }
    printf("parallel parse == serial parse on %ld records\n", checked);
    return 0;
}
