#!/usr/bin/env python
"""Experiment: step-time distribution.  20 device-resident steps timed one by one (CUDA events + host clock), then the
per-kernel event times of 10 more steps -- are outliers GPU-side (a kernel takes longer) or host-side (gaps)?"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
pairs = 10_000_000
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, pairs, 100, seed=SEED)
nreads = 2 * pairs
kc = tg.KmerCounter(ctx, K, True, expected_keys=int(tx_offs[-1]) + int(nreads * 100 * 0.005 * K * 0.68) + (1 << 20))
offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(101)
d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
d1, d2, d3 = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
def step():
    kc.clear(); kc.add_records_dev(d_recs, nbytes); kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
for _ in range(3): step()
ctx.sync()
ev, wall = [], []
for i in range(20):
    t0 = time.perf_counter(); ctx.timer_start(); step(); ms = ctx.timer_stop(); wall.append((time.perf_counter() - t0) * 1e3); ev.append(ms)
print(json.dumps({"event_ms": [round(x, 1) for x in ev], "wall_ms": [round(x, 1) for x in wall]}))
ctx.set("kernel_timing", 1)
rows = []
for i in range(10):
    ctx.kernel_times(); t0 = time.perf_counter(); step(); ctx.sync(); w = (time.perf_counter() - t0) * 1e3
    kt = ctx.kernel_times()
    rows.append({"wall": round(w, 1), **{k: round(v[0], 1) for k, v in kt.items()}})
for r in rows: print(json.dumps(r))
# 20 steps in ONE timed region (what bench.py does)
ctx.set("kernel_timing", 0)
ctx.sync(); ctx.timer_start()
for i in range(20): step()
print(json.dumps({"20_steps_ms_per_step": round(ctx.timer_stop() / 20, 2)}))
