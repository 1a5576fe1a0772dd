#!/usr/bin/env python
"""Experiment: direct vs partitioned count on configs[1]-shaped reads; phase timings, partition-size sweep."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, SEED, K

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--read-len", type=int, default=100)
ap.add_argument("--ntx", type=int, default=20_000)
ap.add_argument("--load", type=float, default=0.45)
ap.add_argument("--parts", default="8,16,32,64")
ap.add_argument("--sigma", type=float, default=2.0)
ap.add_argument("--p1bins", default="")
ap.add_argument("--groups", default="8")
args = ap.parse_args()

ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(args.ntx, SEED, args.sigma)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, args.pairs, args.read_len, seed=SEED)
nreads = 2 * args.pairs
npos = nreads * (args.read_len - K + 1)
expected = int(tx_offs[-1]) + int(nreads * args.read_len * 0.005 * K * 1.15) + (1 << 20)

def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        ctx.sync(); ctx.timer_start(); fn(); ms = ctx.timer_stop(); best = min(best, ms)
    return best

out = []
# direct
ctx.set("count_mode", "direct")
kc = tg.KmerCounter(ctx, K, True, expected_keys=expected)
def direct():
    kc.add_records_dev(d_recs, nbytes)
kc.clear(); ms = timed(direct, 1); info = kc.info()
kc.clear(); ms = min(ms, timed(direct, 1))
out.append({"mode": "direct", "ms": ms, "gkmers_s": npos / ms / 1e6, "geometry": kc.geometry(), **info})
print(json.dumps(out[-1]), flush=True)
distinct = info["distinct"]
kc.close()

for pm in [int(x) for x in args.parts.split(",")]:
    ctx.set("part_mb", pm)
    ctx.set("count_mode", "log")
    kc = tg.KmerCounter(ctx, K, True, expected_keys=int(distinct * 0.45 / args.load))
    for pf in (1, 0):
        ctx.set("replay_prefetch", pf)
        kc.clear(); ms = timed(lambda: kc.add_records_dev(d_recs, nbytes), 1)
        kc.clear(); ms = min(ms, timed(lambda: kc.add_records_dev(d_recs, nbytes), 1))
        inf = kc.info()
        assert inf["distinct"] == distinct, (inf, distinct)
        out.append({"mode": "log", "part_mb": pm, "prefetch": pf, "ms": ms, "gkmers_s": npos / ms / 1e6,
                    "geometry": kc.geometry(), **inf})
        print(json.dumps(out[-1]), flush=True)
    # phases separately through the sharded entry points (same kernels)
    subcap, nparts, p0, nl = kc.geometry()
    nbins = max(nparts, 512)
    cap = tg.sharded.log_capacity(nbytes, nbins)
    keys = ctx.dev_alloc(nbins * cap * 8); cur = ctx.dev_alloc(nbins * 4); hp = ctx.dev_alloc(64); ctx.memset(hp, 0, 64)
    def p1():
        ctx.memset(cur, 0, nbins * 4)
        kc.partition_dev(d_recs, nbytes, nbins, cap, keys, cur, hp)
    ms1 = timed(p1, 2)
    ms2 = {}
    if nbins == nparts:
        for G in [int(x) for x in args.groups.split(",")]:
            ctx.set("replay_groups", G)
            kc.clear()
            m = timed(lambda: kc.replay_log_dev(keys, cur, None, 1, cap), 1)
            kc.clear()
            ms2[G] = round(min(m, timed(lambda: kc.replay_log_dev(keys, cur, None, 1, cap), 1)), 2)
            assert kc.info()["distinct"] == distinct
    curh = ctx.d2h(cur, nbins * 4, np.uint32)
    out.append({"mode": "phases", "part_mb": pm, "phase1_ms": ms1, "phase2_ms": ms2, "entries": int(curh.sum()),
                "entries_per_pos": float(curh.sum()) / npos, "bin_fill_max_over_mean": float(curh.max() / curh.mean())})
    print(json.dumps(out[-1]), flush=True)
    ctx.dev_free(keys); ctx.dev_free(cur)
    kc.close()
for nb in [int(x) for x in args.p1bins.split(",") if x]:
    cap = tg.sharded.log_capacity(nbytes, nb)
    keys = ctx.dev_alloc(nb * cap * 8); cur = ctx.dev_alloc(nb * 4); hp = ctx.dev_alloc(64); ctx.memset(hp, 0, 64)
    kc = tg.KmerCounter(ctx, K, True, expected_keys=1000)
    def p1():
        ctx.memset(cur, 0, nb * 4)
        kc.partition_dev(d_recs, nbytes, nb, cap, keys, cur, hp)
    ms1 = timed(p1, 2)
    out.append({"mode": "phase1_only", "nbins": nb, "phase1_ms": ms1})
    print(json.dumps(out[-1]), flush=True)
    ctx.dev_free(keys); ctx.dev_free(cur); kc.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "exp_count.jsonl"), "w") as f:
    for o in out:
        f.write(json.dumps(o) + "\n")
