#!/usr/bin/env python
"""Experiment: coverage-stats kernel with / without the hot table, hint and size variants."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--sigma", type=float, default=2.0)
ap.add_argument("--variants", default="0:0:60,2097152:0:60,2097152:1:60,2097152:2:60,2097152:3:60,1048576:0:60,4194304:0:60,2097152:0:30")
a = ap.parse_args()
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED, a.sigma)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.pairs, 100, seed=SEED)
nreads = 2 * a.pairs
npos = nreads * 76
kc = tg.KmerCounter(ctx, K, True, expected_keys=int(154e6 * a.pairs / 10e6) + (1 << 20))
kc.add_records_dev(d_recs, nbytes)
print(kc.info(), flush=True)
offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(101)
d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
d1, d2, d3 = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
ref = None
ctx.set("kernel_timing", 1)
for v in a.variants.split(","):
    hk, hints, load = [int(x) for x in v.split(":")]
    ctx.set("hot_keys", hk); ctx.set("hot_hints", hints); ctx.set("hot_load_pct", load)
    p, n = kc.slots_dev()          # drops the hot table: the next stats call rebuilds it with the new settings
    best = 1e9
    for rep in range(2):
        ctx.kernel_times()
        kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
        kt = ctx.kernel_times()
        best = min(best, kt["k_cov_stats"][0])
    med = ctx.d2h(d1, 4 * nreads, np.uint32)
    if ref is None: ref = med
    assert np.array_equal(ref, med)
    print(json.dumps({"hot_keys": hk, "hints": hints, "load": load, "stats_ms": round(best, 2), "gkmers_s": round(npos / best / 1e6, 2),
                      "other": {k: round(x[0], 2) for k, x in kt.items() if k != "k_cov_stats"}}), flush=True)
