#!/usr/bin/env python
"""k = 32 goldens (Inchworm allows kmer_length <= 32, Inchworm/src/KmerCounter.cpp:15-17): reads.fa plus records that hold
the 32-mers a 64-bit key cannot tag -- poly-A / poly-T runs (the all-zero key and its reverse complement), runs broken
after 31 bases, lower case -- through the UNMODIFIED reference binary oracle/_ref/fastaToKmerCoverageStats.
Run in the build container only (needs oracle/_ref):

    python tests/golden/make_golden_k32.py
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
ENV = dict(os.environ, LC_ALL="C", OMP_NUM_THREADS="1")


def main():
    text = open(os.path.join(HERE, "reads.fa"), "rb").read()
    extra = [
        (">polyT/1", b"T" * 50), (">polyA32/1", b"A" * 32), (">polyA33/2", b"a" * 33), (">polyT32/1", b"T" * 32),
        (">a31c/1", b"A" * 31 + b"C" + b"A" * 40), (">t31g/2", b"T" * 31 + b"G" + b"T" * 33),
        (">polyC/1", b"C" * 45), (">polyG/2", b"G" * 45), (">a31/1", b"A" * 31),
        (">mix/1", b"ACGT" * 4 + b"A" * 36 + b"TTTT" + b"A" * 34 + b"N" + b"A" * 32),
    ]
    for n, s in extra:
        text += n.encode() + b"\n" + s + b"\n"
    fa = os.path.join(HERE, "reads_k32.fa")
    open(fa, "wb").write(text)
    stats = os.path.join(REF, "fastaToKmerCoverageStats")
    for mode in ("--DS", "--SS"):
        with open(os.path.join(HERE, f"stats_k32_{mode[2:]}.expected"), "wb") as f:
            subprocess.run([stats, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "32", "--num_threads", "1", mode],
                           check=True, env=ENV, stdout=f, stderr=subprocess.DEVNULL)
    with open(os.path.join(HERE, "stats_k32_capture.expected"), "wb") as f:
        subprocess.run([stats, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "32", "--num_threads", "1",
                        "--capture_coverage_info"], check=True, env=ENV, stdout=f, stderr=subprocess.DEVNULL)
    print("k = 32 goldens written to", HERE)


if __name__ == "__main__":
    main()
