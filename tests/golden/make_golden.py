#!/usr/bin/env python
"""Regenerates tests/golden/*: small seeded inputs plus the outputs of the UNMODIFIED reference binaries
(oracle/_ref/{fastaToKmerCoverageStats,ReadsToTranscripts}, built by oracle/Makefile.ref from /root/reference).
Run in the build container only (needs oracle/_ref); the committed files are what the GPU box checks against.

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import synthdata  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
ENV = dict(os.environ, LC_ALL="C", OMP_NUM_THREADS="1")


def run(cmd, stdout=None):
    subprocess.run(cmd, check=True, env=ENV, stdout=stdout, stderr=subprocess.DEVNULL)


def main():
    rng = np.random.default_rng(424242)
    txs = synthdata.transcriptome(rng, 30, mean_len=700, min_len=200, max_len=2500)
    reads = synthdata.reads_from(rng, txs, 500, 90, lower_rate=0.15, var_len=True)
    names = [">r%d/%d" % (i, 1 + i % 2) for i in range(len(reads))]
    # edge cases: short, exactly k, k+1, Ns, empty, odd headers, chimeras, low complexity
    extra = [
        (">short/1", b"ACGTACGTAC"), (">exactK/1", txs[0][:25]), (">kplus1/1", txs[0][5:31]), (">kplus2/1", txs[0][40:67]),
        (">empty/1", b""), (">allN/1", b"N" * 40), (">polyA/1", b"A" * 60), (">dinuc/1", b"AC" * 40),
        (">name with spaces/1 extra words", txs[1][100:176]), (">tab\tsep/2", txs[2][50:126]),
        (">nmid/1", txs[3][10:40] + b"N" + txs[3][41:86]), (">lower/1", txs[4][20:96].lower()),
        (">chimAB/1", txs[5][:50] + txs[6][:50]), (">chimBA/1", txs[6][:50] + txs[5][:50]),
        (">rc/1", synthdata.revcomp(txs[7][30:106])), (">iupac/1", txs[8][:30] + b"RYKM" + txs[8][34:80]),
        (">long/1", txs[9][:700]),
    ]
    for n, s in extra:
        names.append(n)
        reads.append(s)
    text = synthdata.fasta_text(names, reads)
    # a multi-line record, a blank line inside a record, spaces inside a sequence line
    text += b">multi/1\n" + txs[10][:60] + b"\n" + txs[10][60:120] + b"\n\n" + txs[10][120:150] + b"\n"
    text += b">spaced/2\n" + txs[11][:40] + b" " + txs[11][40:80] + b"\n"
    open(os.path.join(HERE, "reads.fa"), "wb").write(text)
    # same file, but the last record lacks its trailing newline (kept by Inchworm's reader, dropped by Chrysalis')
    open(os.path.join(HERE, "reads_nonl.fa"), "wb").write(text + b">last/1\n" + txs[12][:80])

    bnames, bundles = synthdata.bundles_from(rng, txs, max_contigs=3, share_every=3)
    bundles[2] = bundles[2].lower()
    open(os.path.join(HERE, "bundles.fa"), "wb").write(synthdata.fasta_text(bnames, bundles))

    # a jellyfish-style dump for the --kmers loader: both strands of some k-mers, a bad-length record, a big count
    from oracle import oracle_py as orc
    import trinityrnaseq_b200 as tg
    recs, offs = tg.records_from_sequences([r for r in reads if r])
    keys, cnts = orc.jf_count(recs, 25, True, 2)
    lines = []
    for i, (kk, c) in enumerate(zip(keys, cnts)):
        kmer = tg.packed_to_kmer(kk, 25)
        lines.append(">%d\n%s\n" % (c, kmer))
        if i % 50 == 0:
            lines.append(">3\n%s\n" % synthdata.revcomp(kmer.encode()).decode())
        if i == 10:
            lines.append(">7\nACGTACGTACGTACGTACGTACGTACG\n")          # 27-mer: reported and skipped
        if i == 20:
            lines.append(">4000000000\n%s\n" % tg.packed_to_kmer(keys[21], 25))
    open(os.path.join(HERE, "kmers_L2.fa"), "w").write("".join(lines))

    stats = os.path.join(REF, "fastaToKmerCoverageStats")
    r2t = os.path.join(REF, "ReadsToTranscripts")
    for tag, fa in (("", "reads.fa"), ("_nonl", "reads_nonl.fa")):
        for mode in ("--DS", "--SS"):
            with open(os.path.join(HERE, f"stats{tag}_{mode[2:]}.expected"), "wb") as f:
                run([stats, "--reads", os.path.join(HERE, fa), "--kmers_from_reads", os.path.join(HERE, fa), "--kmer_size",
                     "25", "--num_threads", "1", mode], stdout=f)
    with open(os.path.join(HERE, "stats_capture.expected"), "wb") as f:
        run([stats, "--reads", os.path.join(HERE, "reads.fa"), "--kmers_from_reads", os.path.join(HERE, "reads.fa"),
             "--num_threads", "1", "--capture_coverage_info"], stdout=f)
    with open(os.path.join(HERE, "stats_kmers_L2.expected"), "wb") as f:
        run([stats, "--reads", os.path.join(HERE, "reads.fa"), "--kmers", os.path.join(HERE, "kmers_L2.fa"), "--kmer_size",
             "25", "--num_threads", "1", "--DS"], stdout=f)
    with open(os.path.join(HERE, "stats_k21.expected"), "wb") as f:
        run([stats, "--reads", os.path.join(HERE, "reads.fa"), "--kmers_from_reads", os.path.join(HERE, "reads.fa"),
             "--kmer_size", "21", "--num_threads", "1"], stdout=f)
    for tag, fa in (("", "reads.fa"), ("_nonl", "reads_nonl.fa")):
        for mode, flags in (("ds", []), ("strand", ["-strand"])):
            out = os.path.join(HERE, f"r2t{tag}_{mode}.expected")
            run([r2t, "-i", os.path.join(HERE, fa), "-f", os.path.join(HERE, "bundles.fa"), "-o", out, "-t", "1",
                 "-max_mem_reads", "50000000", "-p", "0"] + flags)
    out = os.path.join(HERE, "r2t_p10_chunk100.expected")
    run([r2t, "-i", os.path.join(HERE, "reads.fa"), "-f", os.path.join(HERE, "bundles.fa"), "-o", out, "-t", "1",
         "-max_mem_reads", "100", "-p", "10"])
    print("golden files written to", HERE)


if __name__ == "__main__":
    main()
