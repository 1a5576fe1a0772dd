#!/usr/bin/env python
"""Experiment: phase 1 with 256 bins and the default 1.2 slack reported an overflow once -- reproduce and look at the cursors."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
pairs = 10_000_000
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, pairs, 100, seed=SEED)
npos = 2 * pairs * 76
kc = tg.KmerCounter(ctx, K, True, expected_keys=1000)
for nb, cap in ((256, tg.sharded.log_capacity(npos, 256)), (256, 7126032 + 16), (512, tg.sharded.log_capacity(npos, 512))):
    keys = ctx.dev_alloc(nb * cap * 8); cur = ctx.dev_alloc(nb * 4); hp = ctx.dev_alloc(64); ctx.memset(hp, 0, 64)
    for rep in range(2):
        ctx.memset(cur, 0, nb * 4)
        kc.partition_dev(d_recs, nbytes, nb, cap, keys, cur, hp)
        err = None
        try:
            ctx.sync()
        except tg.TrinityGpuError as e:
            err = str(e)[:60]
        curh = ctx.d2h(cur, nb * 4, np.uint32).astype(np.int64)
        print(json.dumps({"nbins": nb, "cap": cap, "rep": rep, "err": err, "max_fill": int(curh.max()), "min_fill": int(curh.min()),
                          "sum": int(curh.sum()), "over": int((curh > cap).sum())}), flush=True)
    ctx.dev_free(keys); ctx.dev_free(cur); ctx.dev_free(hp)
