// CPU check of the host-side FASTA readers of the drop-in executables (trinityrnaseq_b200/host/fasta_io.hpp): prints
// what each reader extracts from a file, one record per line, fields separated by \x01, so that the Python test can
// compare it with the oracle's restatements of the reference readers.
//   readers_dump inchworm|dnastream|bundles FILE
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "fasta_io.hpp"
using namespace tgio;

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FileView fv;
    std::string err;
    if (!fv.open(argv[2], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const std::string mode = argv[1];
    auto emit = [](const char* a, size_t an, const char* b, size_t bn) {
        fwrite(a, 1, an, stdout); fputc(1, stdout); fwrite(b, 1, bn, stdout); fputc('\n', stdout);
    };
    if (mode == "inchworm") {              // the way the stats tool drives it: sequences appended straight into a batch
        InchwormFastaReader rd(fv.data, fv.size);
        RecordBatch rb;
        const char* h; size_t hl;
        while (true) {
            const size_t before = rb.recs.size();
            if (!rd.next(&h, &hl, rb.recs)) break;
            const char* acc; size_t al;
            accession_of(h, hl, &acc, &al);
            emit(acc, al, rb.recs.data() + before, rb.recs.size() - before);
            rb.end_record();
        }
    } else if (mode == "dnastream") {
        DnaStreamReader rd(fv.data, fv.size);
        RecordBatch rb;
        const char* name; size_t nl;
        std::string nm;
        while (true) {
            const size_t before = rb.recs.size();
            if (!rd.next(&name, &nl, rb.recs)) break;
            format_read_name(name, nl, nm);
            std::string both(name, nl);
            both.push_back(2);
            both += nm;
            emit(both.data(), both.size(), rb.recs.data() + before, rb.recs.size() - before);
            rb.end_record();
        }
    } else if (mode == "bundles") {
        RecordBatch rb;
        std::vector<std::string> names;
        read_bundles(fv.data, fv.size, rb, names);
        for (size_t i = 0; i < rb.count(); i++) emit(names[i].data(), names[i].size(), rb.seq(i), rb.seq_len(i));
    } else return 2;
    return 0;
}
