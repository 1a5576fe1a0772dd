#!/usr/bin/env python
"""ncu / timing target: the whole hot path once or twice on configs[1]-shaped reads (count -> dump -L 2 -> statistics on
both tables -> label + assign), with per-kernel CUDA-event times printed as one JSON line.
  python tools/prof_step.py [--pairs N] [--reps R] [--stats-load 0.4] [--no-r2t]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, make_bundles, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--read-len", type=int, default=100)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--stats-load", type=float, default=0.40)
ap.add_argument("--no-r2t", action="store_true")
ap.add_argument("--pin", action="store_true", help="pin the device record buffer: one locus order for all calls")
ap.add_argument("--set", action="append", default=[], help="ctx knob key=value")
ap.add_argument("--count-variants", default="", help="';'-separated lists of knobs (k=v,k=v): the count is timed once per list")
a = ap.parse_args()
ctx = tg.Context(0)
for kv in a.set:
    k_, v_ = kv.split("=")
    ctx.set(k_, v_)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.pairs, a.read_len, seed=SEED)
nreads = 2 * a.pairs
expected = int(tx_offs[-1]) + int(nreads * a.read_len * 0.005 * K * 0.68) + (1 << 20)
kc = tg.KmerCounter(ctx, K, True, expected_keys=expected)
offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(a.read_len + 1)
d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
d1, d2, d3 = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
q = None
out = {}
if a.pin:
    ctx.records_pin_dev(d_recs, d_offs, nreads)
if a.count_variants:
    out["count_variants"] = {}
    for var in a.count_variants.split(";"):
        for kv in var.split(","):
            k_, v_ = kv.split("=")
            ctx.set(k_, v_)
        best = None
        for rep in range(2):
            ctx.set("kernel_timing", 1); ctx.kernel_times()
            kc.clear()
            kc.add_records_dev(d_recs, nbytes)
            ctx.sync()
            kt = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
            if best is None or sum(kt.values()) < sum(best.values()):
                best = kt
        out["count_variants"][var] = best
for rep in range(a.reps):
    ctx.set("kernel_timing", 1); ctx.kernel_times()
    kc.clear()
    kc.add_read_records_dev(d_recs, d_offs, nreads)
    ctx.sync()
    out["count_by_read"] = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    info_by_read = kc.info(); hist_by_read = kc.histo()
    kc.clear()
    if a.pin:
        ctx.locus_prepare_dev(K, recompute=True)
    kc.add_records_dev(d_recs, nbytes)
    ctx.sync()
    out["count"] = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    assert kc.info()["distinct"] == info_by_read["distinct"] and np.array_equal(kc.histo(), hist_by_read), "count by read != logged count"
    if q is None:
        q = kc.compacted(2, load=a.stats_load)
    else:
        kc.compact_into(2, q)
    ctx.sync()
    out["dump_L2"] = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    q.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
    ctx.sync()
    out["stats_min2"] = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    m2 = ctx.d2h(d1, 4 * nreads, np.uint32).copy(); s2 = ctx.d2h(d3, 4 * nreads, np.uint32).copy()
    kc.coverage_stats_dev(d_recs, d_offs, nreads, d1, d2, d3)
    ctx.sync()
    out["stats_full"] = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    assert np.array_equal(m2, ctx.d2h(d1, 4 * nreads, np.uint32)) and np.array_equal(s2, ctx.d2h(d3, 4 * nreads, np.uint32))
out["tables"] = {"count": kc.info(), "min2": q.info()}
if not a.no_r2t:
    brecs, boffs, ncontigs = make_bundles(tx, tx_offs, SEED + 1)
    nb = len(boffs) - 1
    d_b = ctx.dev_records_alloc(brecs.nbytes); ctx.h2d(d_b, brecs)
    d_bo = ctx.dev_alloc(boffs.nbytes); ctx.h2d(d_bo, boffs)
    bt = tg.BundleKmerTable(ctx, K, expected_keys=int(tx_offs[-1]) + (1 << 20))
    d_lut = ctx.dev_alloc(bt.entropy_ok.nbytes); ctx.h2d(d_lut, bt.entropy_ok)
    for rep in range(a.reps):
        ctx.kernel_times()
        bt.clear()
        bt.label_bundles_dev(d_b, brecs.nbytes, d_bo, nb)
        bt.assign_reads_dev(d_recs, d_offs, nreads, d_lut, d1, d2, strand=False)
        ctx.sync()
        out["r2t"] = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    out["tables"]["labels"] = bt.info()
print(json.dumps(out))
