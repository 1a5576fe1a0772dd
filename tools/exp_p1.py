#!/usr/bin/env python
"""Experiment: phase 1 (k_log_tiles) alone against the number of bins, log on this GPU."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--bins", default="256,512,1024,2048,4096")
a = ap.parse_args()
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.pairs, 100, seed=SEED)
npos = 2 * a.pairs * 76
kc = tg.KmerCounter(ctx, K, True, expected_keys=1000)
for nb in [int(x) for x in a.bins.split(",")]:
    cap = tg.sharded.log_capacity(int(npos * 1.6), nb)
    keys = ctx.dev_alloc(nb * cap * 8); cur = ctx.dev_alloc(nb * 4); hp = ctx.dev_alloc(64); ctx.memset(hp, 0, 64)
    best = 1e9
    for rep in range(3):
        ctx.memset(cur, 0, nb * 4)
        ctx.sync(); ctx.timer_start()
        kc.partition_dev(d_recs, nbytes, nb, cap, keys, cur, hp)
        best = min(best, ctx.timer_stop())
    curh = ctx.d2h(cur, nb * 4, np.uint32).astype(np.float64)
    print(json.dumps({"nbins": nb, "cap": cap, "phase1_ms": round(best, 2), "gkmers_s": round(npos / best / 1e6, 2),
                      "entries": int(curh.sum()), "fill_max_over_mean": round(float(curh.max() / curh.mean()), 3)}), flush=True)
    ctx.dev_free(keys); ctx.dev_free(cur); ctx.dev_free(hp)
