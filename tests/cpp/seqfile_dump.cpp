// CPU check of host/seq_file.hpp (what `jellyfish count` feeds the GPU): prints the record buffer it builds from a file,
// and the number of flushes for a given flush threshold.
//   seqfile_dump FILE FLUSH_BYTES
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "seq_file.hpp"
using namespace tgio;
int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FileView fv;
    std::string err;
    if (!fv.open(argv[1], &err)) return 1;
    std::vector<char> recs, all;
    int flushes = 0;
    parse_sequence_file(fv.data, fv.data + fv.size, recs, (size_t)atol(argv[2]), [&]() {
        if (recs.empty()) return;
        if (recs.back() != '\n') { printf("FLUSH NOT AT A RECORD BOUNDARY\n"); exit(3); }
        all.insert(all.end(), recs.begin(), recs.end());
        recs.clear();
        flushes++;
    });
    printf("%d\n", flushes);
    fwrite(all.data(), 1, all.size(), stdout);
    return 0;
}
