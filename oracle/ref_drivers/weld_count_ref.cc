// Test infrastructure: a driver around the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/Makefile.ref into oracle/_ref/weld_count_ref) that runs exactly the weldmer-counting step of GraphFromFasta:
//   NonRedKmerTable kmers(kk); kmers.SetUp(crossover); DNAStringStreamFast seq; seq.ReadStream(reads); kmers.AddData(seq);
// (Chrysalis/analysis/GraphFromFasta.cc:1263,1415-1424) and prints `GetCount(weldmer, 0)` for every candidate, one per
// line, in input order.  usage: weld_count_ref <kk> <candidates.fa> <reads.fa>
#include <stdio.h>
#include <stdlib.h>
#include "analysis/DNAVector.h"
#include "analysis/NonRedKmerTable.h"

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: weld_count_ref <kk> <candidates.fa> <reads.fa>\n"); return 2; }
    const int kk = atoi(argv[1]);
    vecDNAVector crossover;
    crossover.Read(argv[2], false, false, true, 1000000);
    NonRedKmerTable kmers(kk);
    kmers.SetUp(crossover);
    DNAStringStreamFast seq;
    seq.ReadStream(argv[3]);
    kmers.AddData(seq);
    for (int i = 0; i < (int)crossover.size(); i++) printf("%d\n", kmers.GetCount(crossover[i], 0));
    return 0;
}
