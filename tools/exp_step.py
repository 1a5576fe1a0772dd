#!/usr/bin/env python
"""Experiment: where does the device-resident step go?  Wall clock (host sync on both sides) of clear / count / stats,
next to the per-kernel event times of the same step."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--part-mb", type=int, default=0, help="table partition size (0 = default 16)")
a = ap.parse_args()
ctx = tg.Context(0)
if a.part_mb:
    ctx.set("part_mb", a.part_mb)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.pairs, 100, seed=SEED)
nreads = 2 * a.pairs
err_keys = int(nreads * 100 * 0.005 * K * 0.68)
kc = tg.KmerCounter(ctx, K, True, expected_keys=int(tx_offs[-1]) + err_keys + (1 << 20))
offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(101)
d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
d1, d2, d3 = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)


def wall(fn):
    ctx.sync()
    t0 = time.perf_counter()
    fn()
    ctx.sync()
    return (time.perf_counter() - t0) * 1e3


for rep in range(a.reps):
    ctx.set("kernel_timing", 0)
    t_clear = wall(kc.clear)
    t_count = wall(lambda: kc.add_records_dev(d_recs, nbytes))
    t_stats = wall(lambda: kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3))
    ctx.timer_start()
    kc.clear(); kc.add_records_dev(d_recs, nbytes); kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
    t_step = ctx.timer_stop()
    ctx.set("kernel_timing", 1); ctx.kernel_times()
    kc.clear(); kc.add_records_dev(d_recs, nbytes); kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
    kt = ctx.kernel_times()
    print(json.dumps({"rep": rep, "clear_ms": round(t_clear, 2), "count_ms": round(t_count, 2), "stats_ms": round(t_stats, 2),
                      "step_event_ms": round(t_step, 2), "kernels": {k: round(v[0], 2) for k, v in kt.items()},
                      "info": kc.info(), "geometry": kc.geometry()}), flush=True)
