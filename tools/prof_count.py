#!/usr/bin/env python
"""ncu target: one logged count (+ optionally direct count and stats) on configs[1]-shaped reads."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--mode", default="log")
ap.add_argument("--part-mb", type=int, default=16)
ap.add_argument("--stats", action="store_true")
a = ap.parse_args()
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.pairs, 100, seed=SEED)
nreads = 2 * a.pairs
ctx.set("count_mode", a.mode); ctx.set("part_mb", a.part_mb)
expected = int(154e6 * a.pairs / 10e6) + (1 << 20)
kc = tg.KmerCounter(ctx, K, True, expected_keys=expected)
kc.add_records_dev(d_recs, nbytes)
print(kc.info(), kc.geometry())
if a.stats:
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(101)
    d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
    d1, d2, d3 = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
    kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
    ctx.sync()
