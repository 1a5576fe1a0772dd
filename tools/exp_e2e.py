#!/usr/bin/env python
"""Where does the end-to-end (host-buffer) step spend its time?  Wall-clock of each C-ABI call of bench.py's host_step on
configs[1]-shaped reads: hold, clear, tg_count_reads (upload + count), tg_cov_stats (offsets up, statistics, results down),
release; plus plain H2D / D2H copies of the same sizes for scale.    python tools/exp_e2e.py [--pairs N]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trinityrnaseq_b200 as tg
from trinityrnaseq_b200 import _lib
from bench import make_transcriptome, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--read-len", type=int, default=100)
ap.add_argument("--set", action="append", default=[])
a = ap.parse_args()
ctx = tg.Context(0)
for kv in a.set:
    k_, v_ = kv.split("=")
    ctx.set(k_, v_)
L = _lib.lib()
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.pairs, a.read_len, seed=SEED)
nreads = 2 * a.pairs
recs_host, o0 = ctx.pinned((nbytes,), np.uint8)
ctx.d2h(d_recs, recs_host)
offs_host, o1 = ctx.pinned((nreads + 1,), np.uint64)
offs_host[:] = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(a.read_len + 1)
med, o2 = ctx.pinned((nreads,), np.uint32); mean, o3 = ctx.pinned((nreads,), np.float32); sd, o4 = ctx.pinned((nreads,), np.float32)
expected = int(tx_offs[-1]) + int(nreads * a.read_len * 0.005 * K * 0.68) + (1 << 20)
kc = tg.KmerCounter(ctx, K, True, expected_keys=expected)
kc.set_count_floor(2)
out = []
for rep in range(4):
    t = [time.perf_counter()]
    ctx.records_hold(recs_host); t.append(time.perf_counter())
    kc.clear(); t.append(time.perf_counter())
    _lib.check(L.tg_count_reads(kc._h, recs_host.ctypes.data, nbytes, 1)); t.append(time.perf_counter())
    _lib.check(L.tg_cov_stats(kc._h, recs_host.ctypes.data, offs_host.ctypes.data, nreads, 1, med.ctypes.data, mean.ctypes.data,
                              sd.ctypes.data, None)); t.append(time.perf_counter())
    ctx.records_release(); t.append(time.perf_counter())
    names = ["hold", "clear", "count_reads", "cov_stats", "release"]
    out.append({n: round((t[i + 1] - t[i]) * 1e3, 2) for i, n in enumerate(names)})
    out[-1]["total"] = round((t[-1] - t[0]) * 1e3, 2)
ctx.sync()
t0 = time.perf_counter(); ctx.h2d(d_recs, recs_host); ctx.sync(); t1 = time.perf_counter()
d_o = ctx.dev_alloc(offs_host.nbytes)
ctx.h2d(d_o, offs_host); ctx.sync(); t2 = time.perf_counter()
d_m = ctx.dev_alloc(4 * nreads)
t3 = time.perf_counter(); ctx.d2h(d_m, med); ctx.sync(); t4 = time.perf_counter()
print(json.dumps({"steps": out, "h2d_reads_ms": round((t1 - t0) * 1e3, 2), "h2d_gbs": round(nbytes / (t1 - t0) / 1e9, 1),
                  "h2d_offs_ms": round((t2 - t1) * 1e3, 2), "d2h_4B_per_read_ms": round((t4 - t3) * 1e3, 2)}))
