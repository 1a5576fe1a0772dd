"""Shared input of the weldmer-counting tests (GraphFromFasta): seeded reads + candidate kk-mers."""
import numpy as np

import synthdata as synth


def weld_case(seed=11, kk=48, nreads=1500):
    """reads (with lower case, N's, short and long reads) and candidates: windows cut from the transcripts (present, some
    repeated in the list), windows with one base changed (mostly absent), one with an N (can never match), poly-A."""
    rng = np.random.default_rng(seed)
    txs = synth.transcriptome(rng, 12, mean_len=700, min_len=300, max_len=2000)
    reads = synth.reads_from(rng, txs, nreads, 100, lower_rate=0.15, var_len=True)
    reads += [b"", b"ACGT" * 5, b"A" * 130, txs[0][:kk], txs[0][:kk - 1], txs[1][: 3 * kk].lower(), b"N" * 70, txs[2][:400]]
    cands = []
    for t in txs:
        for p in rng.integers(0, len(t) - kk, 14):
            cands.append(t[p:p + kk])
    for i in range(0, len(cands), 5):                     # single-base variants: absent (unless a read carries that error)
        c = bytearray(cands[i]); j = int(rng.integers(0, kk)); c[j] = ord("A") if c[j] != ord("A") else ord("C")
        cands.append(bytes(c))
    cands += [cands[3], cands[3], cands[40]]              # duplicates share one counter
    cands.append(b"A" * kk)
    withn = bytearray(cands[7]); withn[5] = ord("N"); cands.append(bytes(withn))
    return reads, cands


def write_fasta(path, seqs, prefix="s"):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">" + f"{prefix}{i}".encode() + b"\n" + s + b"\n")
