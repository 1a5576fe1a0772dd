"""CPU, world_size 2 and 3 over gloo: the host-side logic of the hash-sharded table (trinityrnaseq_b200.sharded) --
geometry, read sharding by offset, bin ownership, both exchanges (the equal-split all-to-all and the "peer" exchange
where phase 1 writes into the owners' receive logs -- memory-mapped files stand in for CUDA IPC memory) and their
[src, lp, cap] receive layout, the all-gather into a full replica, the reductions -- driven through a stand-in engine (tests/standin_engine.py; the product
engine is CUDA and is covered by tests/test_gpu_partitioned.py and the multi-GPU bench)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synthdata as synth
from trinityrnaseq_b200 import sharded
from trinityrnaseq_b200.api import records_from_sequences

HERE = os.path.dirname(os.path.abspath(__file__))


def _make_reads():
    rng = np.random.default_rng(77)
    txs = synth.transcriptome(rng, 12, mean_len=500, min_len=150, max_len=1200)
    reads = synth.reads_from(rng, txs, 700, 80, var_len=True)
    reads += [b"A" * 60, b"", b"ACGTN" * 12, txs[0][:25]]
    return reads


def _worker(rank, world, port, out_dir, exchange, coarse):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import oracle_py as orc
    import standin_engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = 25
        reads = _make_reads()
        recs, offs = records_from_sequences(reads)
        ok, oc = orc.jf_count(recs, k, True, 1)
        r0, r1 = sharded.record_range(offs, rank, world)
        mine = recs[int(offs[r0]):int(offs[r1])]
        eng = standin_engine.StandinEngine(k, True, peer_dir=out_dir if exchange == "peer" else None)
        sc = sharded.ShardedKmerCounter(eng, expected_keys_per_rank=len(ok) // world + 64, part_bytes=8 << 10,
                                        max_exchange_bins=world * coarse if coarse else sharded.MAX_EXCHANGE_BINS)
        if coarse:      # k-mers travel in `coarse` bins per rank and are split into partitions by the owner
            assert sc.c == coarse and sc.lp > sc.c and sc.cbins == world * coarse
        else:
            assert sc.c == sc.lp
        assert sc.exchange == exchange          # "auto" picks the peer exchange exactly when the engine offers it
        assert sc.nparts == world * sc.lp and sc.lp >= 2
        assert sc.table.part0 == rank * sc.lp and sc.table.nlocal == sc.lp
        sc.add_records_dev(torch.from_numpy(mine.copy()) if len(mine) else torch.zeros(0, dtype=torch.uint8), len(mine))
        assert sc.size() == len(ok)
        np.testing.assert_array_equal(sc.histo(), orc.jf_histo(oc))
        lk, lc = sc.dump_local()
        # every local key belongs to this rank, and the union over ranks is the global dump
        assert all(sc.owner_of_bin(standin_engine.key_bin(int(x), sc.nparts)) == rank for x in lk)
        np.save(os.path.join(out_dir, f"k{rank}.npy"), lk)
        np.save(os.path.join(out_dir, f"c{rank}.npy"), lc)
        # second batch: counts double, the log buffers are reused
        sc.add_records_dev(torch.from_numpy(mine.copy()) if len(mine) else torch.zeros(0, dtype=torch.uint8), len(mine))
        full = sc.replicate()
        fk, fc = full.dump()
        np.testing.assert_array_equal(fk, ok)
        np.testing.assert_array_equal(fc, 2 * oc)
        # queries on the replica are purely local: every rank gets the single-table answer for its own reads
        okc = orc.KmerCounter(k, True)
        okc.add_records(recs, offs)
        if r1 > r0:
            sub_offs = offs[r0:r1 + 1] - offs[r0]
            gm, gmean, gsd = full.coverage_stats(mine, sub_offs)
            om, omean, osd = okc.coverage_stats(recs, offs)
            # the oracle table holds single counts; the replica doubled them
            okc2 = orc.KmerCounter(k, True)
            for key, c in zip(ok.tolist(), oc.tolist()):
                okc2.add_kmer("".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k)), 2 * c)
            om, omean, osd = okc2.coverage_stats(recs, offs)
            np.testing.assert_array_equal(gm, om[r0:r1])
            np.testing.assert_array_equal(gsd.view(np.uint32), osd[r0:r1].view(np.uint32))
        # the `dump -L 2` replica: only k-mers seen at least twice, yet the same coverage statistics
        rep2 = sc.replicate(min_count=2)
        k2, c2 = rep2.dump()
        keep = (2 * oc) >= 2
        np.testing.assert_array_equal(k2, ok[keep])
        rep3 = sc.replicate(min_count=3)                 # doubled counts: nothing has count 3 exactly, 2 drops out
        k3, c3 = rep3.dump()
        np.testing.assert_array_equal(k3, ok[(2 * oc) >= 3])
        np.testing.assert_array_equal(c3, (2 * oc)[(2 * oc) >= 3])
        rep2 = sc.replicate(min_count=2)
        if r1 > r0:
            gm2, _, gsd2 = rep2.coverage_stats(mine, sub_offs)
            np.testing.assert_array_equal(gm2, om[r0:r1])
            np.testing.assert_array_equal(gsd2.view(np.uint32), osd[r0:r1].view(np.uint32))
        # routed lookups: no replica -- keys to the owners, counts back -- give the same statistics, also under a floor
        n_mine = r1 - r0
        t_recs = torch.from_numpy(mine.copy()) if len(mine) else torch.zeros(0, dtype=torch.uint8)
        sub = (offs[r0:r1 + 1] - offs[r0]) if n_mine else np.zeros(1, np.uint64)
        for min_count, want in ((1, (om, osd)), (3, None)):
            med, mean, sd = np.zeros(max(n_mine, 1), np.uint32), np.zeros(max(n_mine, 1), np.float32), np.zeros(max(n_mine, 1), np.float32)
            sc.coverage_stats_routed_dev(t_recs, len(mine), sub, n_mine, med, mean, sd, min_count=min_count)
            if want is None:                             # what a table rebuilt from `dump -L 3` answers
                okc3 = orc.KmerCounter(k, True)
                for key, c in zip(ok.tolist(), oc.tolist()):
                    if 2 * c >= 3:
                        okc3.add_kmer("".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k)), 2 * c)
                m3, _, s3 = okc3.coverage_stats(recs, offs)
                want = (m3, s3)
            if n_mine:
                np.testing.assert_array_equal(med[:n_mine], want[0][r0:r1])
                np.testing.assert_array_equal(sd[:n_mine].view(np.uint32), want[1][r0:r1].view(np.uint32))
        sc.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,exchange,coarse", [(2, "collective", 0), (3, "collective", 2), (2, "peer", 1), (3, "peer", 0)])
def test_sharded_counter_over_gloo(world, exchange, coarse, tmp_path):
    port = 29500 + (os.getpid() % 400) + world + (10 if exchange == "peer" else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), exchange, coarse), nprocs=world, join=True)
    from oracle import oracle_py as orc
    recs, offs = records_from_sequences(_make_reads())
    ok, oc = orc.jf_count(recs, 25, True, 1)
    ks = np.concatenate([np.load(tmp_path / f"k{r}.npy") for r in range(world)])
    cs = np.concatenate([np.load(tmp_path / f"c{r}.npy") for r in range(world)])
    order = np.argsort(ks, kind="stable")
    np.testing.assert_array_equal(ks[order], ok)
    np.testing.assert_array_equal(cs[order], oc)


def _skew_worker(rank, world, port, out_dir):
    """one k-mer carried by a large share of the reads: its log bin overflows the even-spread head-room, the ranks agree,
    double the head-room and repeat the batch (the sharded path used to abort here)"""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import oracle_py as orc
    import standin_engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = 25
        rng = np.random.default_rng(5)
        txs = synth.transcriptome(rng, 6, mean_len=400, min_len=150, max_len=800)
        hot = b"AGATCGGAAGAGCACACGTCTGAAC"                       # exactly one 25-mer
        # a few thousand spread-out windows plus 8000 copies of one k-mer: with the exact window count as the bound a bin
        # has room for ~1500 entries, the hot k-mer alone brings 4000 per rank
        reads = synth.reads_from(rng, txs, 120, 60, var_len=True) + [hot] * 8000
        recs, offs = records_from_sequences(reads)
        ok, oc = orc.jf_count(recs, k, True, 1)
        assert oc.max() >= 3000
        r0, r1 = sharded.record_range(offs, rank, world)
        mine = recs[int(offs[r0]):int(offs[r1])]
        eng = standin_engine.StandinEngine(k, True, peer_dir=out_dir)
        sc = sharded.ShardedKmerCounter(eng, expected_keys_per_rank=len(ok) // world + 64, part_bytes=2 << 10, max_exchange_bins=32)
        assert sc.lp > sc.c                               # coarse exchange bins + owner-side refine: both can overflow
        nwin = sum(max(0, len(r) - k + 1) for r in reads[r0:r1])
        sc.add_records_dev(torch.from_numpy(mine.copy()), len(mine), max_windows=nwin)
        assert sc.overflow_retries >= 1 and sc._grow >= 2
        assert sc.size() == len(ok)
        np.testing.assert_array_equal(sc.histo(), orc.jf_histo(oc))
        retries = sc.overflow_retries
        sc.add_records_dev(torch.from_numpy(mine.copy()), len(mine), max_windows=nwin)       # the head-room is remembered
        assert sc.overflow_retries == retries
        fk, fc = sc.replicate().dump()
        np.testing.assert_array_equal(fk, ok)
        np.testing.assert_array_equal(fc, 2 * oc)
        sc.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_counter_survives_a_hot_kmer(tmp_path):
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_skew_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_record_range_partitions_every_record_once():
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 200, 1000)
    offs = np.zeros(1001, dtype=np.uint64)
    offs[1:] = np.cumsum(lens + 1)
    for world in (1, 2, 3, 8):
        rr = [sharded.record_range(offs, r, world) for r in range(world)]
        assert rr[0][0] == 0 and rr[-1][1] == 1000
        assert all(rr[i][1] == rr[i + 1][0] for i in range(world - 1))
        sizes = [int(offs[b] - offs[a]) for a, b in rr]
        assert max(sizes) - min(sizes) <= 2 * 201


def test_exchange_bins():
    for world in (1, 2, 3, 4, 8):
        for lp in (1, 2, 64, 512):
            c = sharded.exchange_bins(world, lp)
            assert c >= 1 and lp % c == 0 and c & (c - 1) == 0
            assert world * c <= sharded.MAX_EXCHANGE_BINS or c == 1
            assert c == lp or world * c * 2 > sharded.MAX_EXCHANGE_BINS
    assert sharded.exchange_bins(8, 512) == 16 and sharded.exchange_bins(2, 512) == 64


def test_shard_geometry():
    for world in (1, 2, 4, 8):
        subcap, nparts, lp = sharded.shard_geometry(world, 150_000_000)
        assert nparts == world * lp and nparts <= sharded.MAX_BINS
        assert subcap * lp >= 150_000_000 / sharded.TARGET_LOAD
        assert subcap * 16 <= (16 << 20) or nparts * 2 > sharded.MAX_BINS
