#!/usr/bin/env python
"""k = 32 runs on the direct-insert / CTA-per-read kernels: what does that cost next to k = 31 on the same reads?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
npairs = 2_000_000
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, npairs, 100, seed=SEED)
recs = ctx.d2h(d_recs, nbytes, np.uint8).copy()
n = 2 * npairs
offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(101))
for k in (31, 32):
    with tg.KmerCounter(ctx, k, is_ds=True, expected_keys=60_000_000) as kc:
        for rep in range(2):
            kc.clear()
            ctx.sync(); t0 = time.perf_counter()
            kc.add_records(recs)
            ctx.sync(); t1 = time.perf_counter()
            m, mean, sd = kc.coverage_stats(recs, offs)
            ctx.sync(); t2 = time.perf_counter()
        pos = n * (100 - k + 1)
        print(f"k={k}: count {1e3 * (t1 - t0):.1f} ms, stats {1e3 * (t2 - t1):.1f} ms (host-buffer entry points, {n} reads) -> "
              f"{2 * pos / (t2 - t0) / 1e9:.2f} G positions/s; distinct {kc.size()}, median of medians {int(np.median(m))}")
