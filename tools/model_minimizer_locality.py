#!/usr/bin/env python
"""CPU model (numpy, no GPU): how many DRAM granules does one read's worth of lookups touch under the current table
layout, and under a layout that places a k-mer by its MINIMIZER?

Current layout: bucket = hash(k-mer): every window of a read lands in its own 64-B bucket, so a 100-bp read (76 windows)
touches ~76 granules; ncu measures ~100 B of DRAM traffic per lookup.
Minimizer layout: neighbourhood = hash(canonical minimizer of the k-mer) (w = k - m + 1 consecutive windows share it on
average (w + 1) / 2 times), position inside the neighbourhood = hash(k-mer).  The model counts, per read, the distinct
64-B and 128-B granules its windows map to, and the load a neighbourhood must absorb (distinct k-mers per minimizer),
for the full table (error k-mers included) and for the `dump -L 2` table (what the normalisation pipeline queries).

    python tools/model_minimizer_locality.py [--ntx 1500] [--reads 20000] [--m 13]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K = 25
MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd)
    x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53)
    x ^= x >> np.uint64(33)
    return x


def pack_windows(codes, k):
    """codes: uint8 array of 0..3 (4 = invalid).  -> (packed forward k-mers, valid mask) for every window"""
    n = len(codes) - k + 1
    fwd = np.zeros(n, dtype=np.uint64)
    bad = np.zeros(n, dtype=np.int32)
    for i in range(k):
        c = codes[i:i + n]
        fwd = (fwd << np.uint64(2)) | (c & 3).astype(np.uint64)
        bad += (c > 3)
    return fwd, bad == 0


def revcomp_packed(v, k):
    r = np.zeros_like(v)
    x = v.copy()
    for _ in range(k):
        r = (r << np.uint64(2)) | (np.uint64(3) - (x & np.uint64(3)))
        x >>= np.uint64(2)
    return r


def canonical(v, k):
    return np.minimum(v, revcomp_packed(v, k))


def minimizers(codes, k, m):
    """canonical minimizer (smallest hash of a canonical m-mer) of every k-window"""
    mm, okm = pack_windows(codes, m)
    hm = mix(canonical(mm, m))
    hm[~okm] = MASK64
    w = k - m + 1
    n = len(codes) - k + 1
    out = np.full(n, MASK64, dtype=np.uint64)
    for j in range(w):
        out = np.minimum(out, hm[j:j + n])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntx", type=int, default=1500)
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--m", type=int, default=13)
    ap.add_argument("--err", type=float, default=0.005)
    a = ap.parse_args()
    rng = np.random.default_rng(20251017)
    lens = np.clip(np.round(rng.lognormal(np.log(1500), 0.6, a.ntx)), 300, 10000).astype(int)
    txs = [rng.integers(0, 4, L).astype(np.uint8) for L in lens]
    weights = rng.lognormal(0, 2.0, a.ntx) * lens
    weights /= weights.sum()
    # reads: random position, random strand, substitution errors
    reads = []
    for t in rng.choice(a.ntx, a.reads, p=weights):
        tx = txs[t]
        p = rng.integers(0, len(tx) - a.read_len + 1)
        r = tx[p:p + a.read_len].copy()
        if rng.integers(0, 2):
            r = (3 - r)[::-1]
        e = rng.random(a.read_len) < a.err
        r[e] = (r[e] + rng.integers(1, 4, int(e.sum()))) & 3
        reads.append(r)
    # all windows of all reads
    keys, mins, rid = [], [], []
    for i, r in enumerate(reads):
        f, ok = pack_windows(r, K)
        keys.append(canonical(f, K)[ok])
        mins.append(minimizers(r, K, a.m)[ok])
        rid.append(np.full(int(ok.sum()), i))
    keys, mins, rid = np.concatenate(keys), np.concatenate(mins), np.concatenate(rid)
    uniq, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    out = {"reads": a.reads, "windows": int(len(keys)), "distinct_kmers": int(len(uniq)),
           "singletons": int((cnt == 1).sum()), "m": a.m}

    def granules_per_read(bucket_of_window, slots_per_granule):
        g = bucket_of_window // np.uint64(slots_per_granule)
        pair = np.unique(np.stack([rid.astype(np.uint64), g]), axis=1)
        return pair.shape[1] / a.reads

    for table, keep in (("full", np.ones(len(uniq), bool)), ("dump_L2", cnt >= 2)):
        nkeys = int(keep.sum())
        res = {"keys": nkeys}
        # current layout: slot = hash(k-mer) over a table at load 0.4 (home slot only; probing adds a little)
        nslots = np.uint64(int(nkeys / 0.4))
        slot = mix(keys) % nslots
        res["current"] = {"granules64_per_read": round(granules_per_read(slot, 4), 1),
                          "granules128_per_read": round(granules_per_read(slot, 8), 1)}
        # minimizer layout: neighbourhood of N slots by hash(minimizer), slot inside by hash(k-mer)
        kept = keep[inv]
        per_min = np.unique(np.stack([mins[kept], keys[kept]]), axis=1)          # distinct (minimizer, k-mer)
        _, load = np.unique(per_min[0], return_counts=True)
        res["kmers_per_minimizer"] = {"mean": round(float(load.mean()), 1), "p50": int(np.percentile(load, 50)),
                                      "p99": int(np.percentile(load, 99)), "max": int(load.max())}
        for nb_slots in (8, 16, 32, 64):
            nnb = np.uint64(max(int(nkeys / 0.4 / nb_slots), 1))
            nb = mix(mins ^ np.uint64(0x5bd1e995)) % nnb
            inside = mix(keys) % np.uint64(nb_slots)
            slot = nb * np.uint64(nb_slots) + inside
            # share of neighbourhoods whose distinct k-mers exceed their slots (they spill into the next one)
            nbk = np.unique(np.stack([mix(per_min[0] ^ np.uint64(0x5bd1e995)) % nnb, per_min[1]]), axis=1)
            _, occ = np.unique(nbk[0], return_counts=True)
            res[f"minimizer_nb{nb_slots}"] = {
                "granules64_per_read": round(granules_per_read(slot, 4), 1),
                "granules128_per_read": round(granules_per_read(slot, 8), 1),
                "neighbourhoods_over_capacity_pct": round(100.0 * float((occ > nb_slots).sum()) / len(occ), 2),
                "keys_in_overfull_pct": round(100.0 * float(occ[occ > nb_slots].sum()) / float(occ.sum()), 2)}
        out[table] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
