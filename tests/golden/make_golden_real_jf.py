#!/usr/bin/env python
"""Adds the fixtures that pin the jellyfish side (stage J) to REAL jellyfish output as far as the reference tree allows:
the reference ships one real dump, trinity_ext_sample_data/test_Inchworm/jellyfish.kmers.fa.gz (55,289 records; the
reads it was counted from are not in the tree).  Written here:

  real_jf_dump_head.fa       its first 3000 records, verbatim (>COUNT\\nKMER\\n in jellyfish's own hash order)
  real_jf_reads.fa           reads stitched from those k-mers (record i = k-mer i + k-mer i+1, some reverse-complemented,
                             one with an N), so that their coverage is decided by the counts in the dump
  stats_real_jf.expected     the UNMODIFIED reference tool: fastaToKmerCoverageStats --kmers real_jf_dump_head.fa
                             --reads real_jf_reads.fa --DS (oracle/_ref, single thread)

Run in the build container only (needs /root/reference and oracle/_ref):   python tests/golden/make_golden_real_jf.py
"""
import gzip
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = "/root/reference/trinity_ext_sample_data/test_Inchworm/jellyfish.kmers.fa.gz"
REF = os.path.join(ROOT, "oracle", "_ref", "fastaToKmerCoverageStats")
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def main():
    lines = gzip.open(SRC, "rb").read().split(b"\n")
    head = lines[:6000]
    open(os.path.join(HERE, "real_jf_dump_head.fa"), "wb").write(b"\n".join(head) + b"\n")
    kmers = head[1::2]
    out = []
    for i in range(0, len(kmers) - 1, 3):
        s = kmers[i] + kmers[i + 1]
        if i % 2:
            s = s.translate(COMP)[::-1]
        if i % 50 == 0:
            s = s[:30] + b"N" + s[31:]
        out.append(b">jf%d/1\n%s\n" % (i, s))
    open(os.path.join(HERE, "real_jf_reads.fa"), "wb").write(b"".join(out))
    with open(os.path.join(HERE, "stats_real_jf.expected"), "wb") as f:
        subprocess.run([REF, "--reads", os.path.join(HERE, "real_jf_reads.fa"), "--kmers",
                        os.path.join(HERE, "real_jf_dump_head.fa"), "--kmer_size", "25", "--num_threads", "1", "--DS"],
                       check=True, stdout=f, stderr=subprocess.DEVNULL, env=dict(os.environ, LC_ALL="C", OMP_NUM_THREADS="1"))
    print("written")


if __name__ == "__main__":
    main()
