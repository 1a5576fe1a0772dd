// seqtk-trinity -- drop-in for the one call Trinity makes of its modified seqtk (read prep):
//
//     cat reads.fq | seqtk-trinity seq -A -R <1|2> [-r] -  >> left.fa            (Trinity:2784-2806)
//
// Reference: trinity-plugins/seqtk-trinity/seqtk.c (stk_seq :396-573, update_read_name_for_Trinity :201-291) over
// kseq.h (kseq_read :170-213, ks_getuntil2 :91-142).  FASTA/FASTQ in (file, '-' = stdin, gzip through zlib like the
// reference's gzopen), single-line FASTA out with the read name normalised to name/1 or name/2; the same exit codes the
// reference's own tests check (testing/test_seqtk_trinity.py): 2 wrong or missing read type, 3 quality and sequence of
// different lengths, 4 an empty sequence, 5 nothing parsed.  CPU only -- this stage is a text filter in front of the
// k-mer path; what it buys is that the filter keeps up with the pipe that feeds it (one pass with memchr over a 4 MiB
// window, one write per 4 MiB of output, instead of a 16 KiB kstream and putchar/puts per record).
//
// Supported options: -A/-a -C -r -U -S -N -1 -2 -L -l -q -Q -X -n -R (all of `seq` that needs no BED file and no sampling);
// -M -c -s -f -V are refused (exit 1) rather than silently ignored.
#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <string>
#include <vector>

namespace {

// ---- input: kstream_t restated over a big window (kseq.h:40-142) --------------------------------------------------
struct Stream {
    gzFile f = nullptr;
    std::vector<unsigned char> buf;
    size_t begin = 0, end = 0;
    bool is_eof = false;
    explicit Stream(gzFile g) : f(g), buf(4u << 20) {}
    bool refill() {                       // false at end of input
        if (is_eof) return false;
        begin = 0;
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        end = n > 0 ? (size_t)n : 0;
        if (end == 0) { is_eof = true; return false; }
        return true;
    }
    int getc() {                          // ks_getc
        if (begin >= end && !refill()) return -1;
        return buf[begin++];
    }
    // ks_getuntil2: bytes up to the delimiter (line: '\n'; space: any isspace) appended to / replacing str; *dret = the
    // delimiter found (0 at end of input).  Returns the length, or -1 when nothing was read and the input is exhausted.
    // A line that ends in '\r' loses it (only when more than one byte long: kseq.h:134).
    long getuntil(bool line, std::string& str, int* dret, bool append) {
        bool gotany = false;
        if (dret) *dret = 0;
        if (!append) str.clear();
        for (;;) {
            if (begin >= end && !refill()) break;
            size_t i;
            if (line) {
                const void* nl = memchr(buf.data() + begin, '\n', end - begin);
                i = nl ? (size_t)((const unsigned char*)nl - buf.data()) : end;
            } else {
                for (i = begin; i < end; ++i) if (isspace(buf[i])) break;
            }
            gotany = true;
            str.append((const char*)buf.data() + begin, i - begin);
            begin = i + 1;
            if (i < end) { if (dret) *dret = buf[i]; break; }
        }
        if (!gotany && is_eof && begin >= end) return -1;
        if (line && str.size() > 1 && str.back() == '\r') str.pop_back();
        return (long)str.size();
    }
};

// ---- one record: kseq_read (kseq.h:170-213) ----------------------------------------------------------------------------
struct Record {
    std::string name, comment, seq, qual;
    bool has_comment_buffer = false;      // the reference passes comment.s == NULL until a header line has carried a comment
    int last_char = 0;
};
// >= 0 sequence length, -1 end of input, -2 quality string truncated / of another length
long read_record(Stream& ks, Record& r) {
    int c;
    if (r.last_char == 0) {               // jump to the next header character, wherever it stands
        while ((c = ks.getc()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        r.last_char = c;
    }
    r.comment.clear(); r.seq.clear(); r.qual.clear();
    if (ks.getuntil(false, r.name, &c, false) < 0) return -1;
    if (c != '\n') { ks.getuntil(true, r.comment, nullptr, false); r.has_comment_buffer = true; }
    while ((c = ks.getc()) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;          // empty line
        r.seq.push_back((char)c);
        ks.getuntil(true, r.seq, nullptr, true);
    }
    if (c == '>' || c == '@') r.last_char = c;
    if (c != '+') return (long)r.seq.size();                  // FASTA
    while ((c = ks.getc()) != -1 && c != '\n') {}             // rest of the '+' line
    if (c == -1) return -2;
    while (ks.getuntil(true, r.qual, nullptr, true) >= 0 && r.qual.size() < r.seq.size()) {}
    r.last_char = 0;
    if (r.seq.size() != r.qual.size()) return -2;
    return (long)r.seq.size();
}

struct Out {
    std::vector<char> buf;
    bool failed = false;
    Out() { buf.reserve(4u << 20); }
    void flush() {
        const char* p = buf.data(); size_t n = buf.size();
        while (n && !failed) {
            const ssize_t w = ::write(1, p, n);
            if (w < 0) { if (errno == EINTR) continue; failed = true; break; }
            p += w; n -= (size_t)w;
        }
        buf.clear();
    }
    void put(const char* p, size_t n) { if (buf.size() + n > (4u << 20)) flush(); buf.insert(buf.end(), p, p + n); }
    void putc(char c) { put(&c, 1); }
};
Out out;
[[noreturn]] void leave(int code) { out.flush(); exit(code); }

// ---- the Trinity modification: name -> name/1 | name/2 (seqtk.c:201-291) ---------------------------------------------
// old format (name ends in /1 or /2): kept, the digit must be the expected one; `_forward` / `_reverse` in the name: cut
// there and tagged; new Illumina format (comment starts with "1:" / "2:"): that digit is checked and becomes the tag;
// anything else: tagged with the expected read type.  A contradiction is exit 2.
void trinity_name(std::string& name, const Record& r, size_t comment_len, unsigned read_type) {
    size_t name_len = name.size();
    const char last = name_len ? name[name_len - 1] : 0;
    size_t found = name.find("_forward");
    if (found == std::string::npos) found = name.find("_reverse");
    const bool cut = found != std::string::npos;
    if (cut) name_len = found;
    auto mismatch = [&](char ch) {
        fprintf(stderr, "Error, found read_type %c but expecting read_type %i\n", ch, (int)read_type);
        leave(2);                         // (what was printed so far -- this record's '>' included -- still goes out, as exit() flushes stdio)
    };
    if (!cut && name_len >= 2 && name[name_len - 2] == '/' && (last == '1' || last == '2')) {
        if ((unsigned)(last - '0') != read_type) mismatch(last);
        return;
    }
    if (r.has_comment_buffer && !cut && comment_len > 1 && r.comment[1] == ':' && (r.comment[0] == '1' || r.comment[0] == '2')) {
        if ((unsigned)(r.comment[0] - '0') != read_type) mismatch(r.comment[0]);
        name.resize(name_len);
        name.push_back('/'); name.push_back(r.comment[0]);
        return;
    }
    name.resize(name_len);
    name.push_back('/'); name.push_back(read_type == 1 ? '1' : '2');
}

// complement of every IUPAC letter, case kept, everything else unchanged (seqtk.c comp_tab :163-180)
unsigned char COMP[256];
void init_comp() {
    for (int i = 0; i < 256; i++) COMP[i] = (unsigned char)i;
    const char* from = "ABCDGHKMRSTVWYN";
    const char* to   = "TVGHCDMKYSABWRN";
    for (int i = 0; from[i]; i++) { COMP[(unsigned char)from[i]] = (unsigned char)to[i]; COMP[(unsigned char)tolower(from[i])] = (unsigned char)tolower(to[i]); }
    COMP[(unsigned char)'U'] = 'A'; COMP[(unsigned char)'u'] = 'a';
    COMP[96] = 64;                        // (the table's lower-case row starts with 64, not 96)
}

int usage_seq(unsigned line_len, int qual_shift) {
    fprintf(stderr, "\nUsage:   seqtk seq [options] <in.fq>|<in.fa>\n\n"
                    "Options: -q INT    mask bases with quality lower than INT [0]\n"
                    "         -X INT    mask bases with quality higher than INT [255]\n"
                    "         -n CHAR   masked bases converted to CHAR; 0 for lowercase [0]\n"
                    "         -l INT    number of residues per line; 0 for 2^32-1 [%d]\n"
                    "         -Q INT    quality shift: ASCII-INT gives base quality [%d]\n"
                    "         -L INT    drop sequences with length shorter than INT [0]\n"
                    "         -r        reverse complement\n"
                    "         -A        force FASTA output (discard quality)\n"
                    "         -C        drop comments at the header lines\n"
                    "         -N        drop sequences containing ambiguous bases\n"
                    "         -1        output the 2n-1 reads only\n"
                    "         -2        output the 2n reads only\n"
                    "         -U        convert all bases to uppercases\n"
                    "         -S        strip of white spaces in sequences\n"
                    "         -R        read_type 1 (left) or 2 (right).  ie. -R 1 or -R 2\n\n", (int)line_len, qual_shift);
    return 1;
}

int seq_main(int argc, char** argv) {
    int c, qual_thres = 0, flag = 0, qual_shift = 33, mask_chr = 0, min_len = 0, max_q = 255;
    unsigned line_len = 0, read_type = 0;
    while ((c = getopt(argc, argv, "N12q:l:Q:aACrn:s:f:M:L:cVUX:SR:")) >= 0) {
        switch (c) {
            case 'a': case 'A': flag |= 1; break;
            case 'C': flag |= 2; break;
            case 'r': flag |= 4; break;
            case '1': flag |= 16; break;
            case '2': flag |= 32; break;
            case 'N': flag |= 128; break;
            case 'U': flag |= 256; break;
            case 'S': flag |= 512; break;
            case 'n': mask_chr = *optarg; break;
            case 'Q': qual_shift = atoi(optarg); break;
            case 'q': qual_thres = atoi(optarg); break;
            case 'X': max_q = atoi(optarg); break;
            case 'l': line_len = (unsigned)atoi(optarg); break;
            case 'L': min_len = atoi(optarg); break;
            case 'R': read_type = (unsigned)atoi(optarg); break;
            case 'M': case 'c': case 's': case 'f': case 'V':
                fprintf(stderr, "seqtk-trinity: option -%c is not supported by this build (Trinity calls `seq -A -R <1|2> [-r]`)\n", c);
                return 1;
            default: break;
        }
    }
    if (argc == optind && isatty(fileno(stdin))) return usage_seq(line_len, qual_shift);
    if (read_type < 1 || read_type > 2) {
        fprintf(stderr, "Error, must specify read type via -R as 1 or 2   ");
        exit(2);
    }
    const bool from_file = optind < argc && strcmp(argv[optind], "-") != 0;
    const char* filename = from_file ? argv[optind] : "-";
    if (line_len == 0) line_len = UINT_MAX;
    gzFile fp = from_file ? gzopen(argv[optind], "r") : gzdopen(fileno(stdin), "r");
    if (fp == nullptr) {
        fprintf(stderr, "[E::%s] failed to open the input file/stream.\n", "stk_seq");
        return 1;
    }
    gzbuffer(fp, 1u << 20);
    Stream ks(fp);
    Record r;
    init_comp();
    qual_thres += qual_shift;
    long n_seqs = 0;
    std::string name;
    for (;;) {
        const long ret = read_record(ks, r);
        if (ret < -1) {
            fprintf(stderr, "Error encountered just after sequence entry[%li]: %s, quals and seq lines dont match in length:\n\n... corrupt file?",
                    n_seqs + 1, r.name.c_str());
            leave(3);
        } else if (ret == -1) {
            break;
        } else if (ret == 0) {
            fprintf(stderr, "Error encountered at sequence entry[%li] ... corrupt file?", n_seqs);
            leave(4);
        }
        ++n_seqs;
        if ((long)r.seq.size() < (long)min_len) continue;
        if (flag & 48) {
            if ((flag & 16) && (n_seqs & 1) == 0) continue;
            if ((flag & 32) && (n_seqs & 1) == 1) continue;
        }
        if (flag & 512) {                 // -S: white space squeezed out of the sequence (and its qualities)
            size_t k = 0;
            if (!r.qual.empty()) {
                for (size_t i = 0; i < r.seq.size(); ++i) if (!isspace((unsigned char)r.seq[i])) r.qual[k++] = r.qual[i];
                r.qual.resize(k);
            }
            k = 0;
            for (size_t i = 0; i < r.seq.size(); ++i) if (!isspace((unsigned char)r.seq[i])) r.seq[k++] = r.seq[i];
            r.seq.resize(k);
        }
        if (!r.qual.empty() && qual_thres > qual_shift) {
            for (size_t i = 0; i < r.seq.size(); ++i)
                if (r.qual[i] < qual_thres || r.qual[i] > max_q) r.seq[i] = mask_chr ? (char)mask_chr : (char)tolower((unsigned char)r.seq[i]);
        }
        if (flag & 256) for (char& ch : r.seq) ch = (char)toupper((unsigned char)ch);
        if (flag & 1) r.qual.clear();
        size_t comment_len = r.comment.size();
        if (flag & 2) comment_len = 0;    // -C: the comment is gone before the name is looked at
        if (flag & 4) {
            const size_t L = r.seq.size();
            for (size_t i = 0; i < L / 2; ++i) {
                const unsigned char c0 = COMP[(unsigned char)r.seq[i]], c1 = COMP[(unsigned char)r.seq[L - 1 - i]];
                r.seq[i] = (char)c1; r.seq[L - 1 - i] = (char)c0;
            }
            if (L & 1) r.seq[L / 2] = (char)COMP[(unsigned char)r.seq[L / 2]];
            if (!r.qual.empty()) for (size_t i = 0; i < L / 2; ++i) std::swap(r.qual[i], r.qual[r.qual.size() - 1 - i]);
        }
        if (flag & 128) {                 // -N: sequences with anything but ACGT (either case) are dropped
            size_t i = 0;
            for (; i < r.seq.size(); ++i) {
                const char u = (char)toupper((unsigned char)r.seq[i]);
                if (u != 'A' && u != 'C' && u != 'G' && u != 'T') break;
            }
            if (i < r.seq.size()) continue;
        }
        // stk_printseq_renamed (:296-323): '@' or '>', the Trinity name, no comment, the sequence (and qualities) in lines
        out.putc(r.qual.empty() ? '>' : '@');
        name = r.name;
        trinity_name(name, r, comment_len, read_type);
        out.put(name.data(), name.size());
        auto put_lines = [&](const std::string& s) {
            if (line_len != UINT_MAX && line_len < s.size()) {
                for (size_t i = 0; i < s.size(); i += line_len) {
                    out.putc('\n');
                    out.put(s.data() + i, std::min<size_t>(line_len, s.size() - i));
                }
                out.putc('\n');
            } else {
                out.putc('\n');
                out.put(s.data(), s.size());
                out.putc('\n');
            }
        };
        put_lines(r.seq);
        if (!r.qual.empty()) { out.putc('+'); put_lines(r.qual); }
    }
    if (n_seqs <= 0) {
        fprintf(stderr, "Error, no records were correctly parsed from %s", filename);
        leave(5);
    }
    gzclose(fp);
    out.flush();
    if (out.failed) { fprintf(stderr, "seqtk-trinity: write failed: %s\n", strerror(errno)); return 1; }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc == 1) {
        fprintf(stderr, "\nUsage:   seqtk <command> <arguments>\nVersion: 1.2-r95-dirty\n\nCommand: seq       common transformation of FASTA/Q\n\n");
        return 1;
    }
    if (strcmp(argv[1], "seq") == 0) {
        seq_main(argc - 1, argv + 1);     // (the reference ignores stk_seq's return value too: usage and open failures leave with 0)
        return 0;
    }
    fprintf(stderr, "[main] unrecognized command '%s'. Abort!\n", argv[1]);
    return 1;
}
