"""GPU parity of the partitioned (logged) count path and of the sharded table, against the CPU oracle (bit-exact).

The logged path only pays on inputs far larger than a unit test, so the tests force it (count_mode=log) and shrink
its blocking units (partition bytes, log bytes, batch bytes) until a few thousand reads exercise every branch:
many partitions, several log flushes, bin overflow -> direct insert, growth by the sampled distinct estimate,
homopolymer run folding, and the N-rank exchange emulated on one GPU by slicing the logs exactly as the
equal-split all-to-all does.
"""
import ctypes as C

import numpy as np
import pytest

import trinityrnaseq_b200 as tg
from trinityrnaseq_b200 import _lib
import synthdata as synth

pytestmark = pytest.mark.gpu

LOG_ENTRY = _lib.lib().tg_log_entry_bytes()      # bytes of one k-mer log entry (key + packed home)


@pytest.fixture()
def ctx():
    c = tg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(424242)
    txs = synth.transcriptome(rng, 50, mean_len=900, min_len=200, max_len=4000)
    reads = synth.reads_from(rng, txs, 8000, 100, lower_rate=0.1, var_len=True)
    reads += [b"A" * 120, b"T" * 77, b"A" * 25, b"ACACACACACACACACACACACACACACACACACACACAC" * 3, b"C" * 33 + b"N" + b"C" * 64,
              b"", b"ACGT", b"N" * 60, txs[0][:25], txs[1][:26], b"acgtnACGTN" * 13] * 3
    return txs, reads


def _dev_records(ctx, recs):
    d = ctx.dev_records_alloc(recs.nbytes)
    ctx.h2d(d, recs)
    return d


@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("k", [25, 31, 17])
def test_logged_count_host_path(ctx, oracle, data, canonical, k):
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, k, canonical, 1)
    ctx.set("count_mode", "log")
    ctx.set("part_bytes", 64 << 10)          # 4096-slot partitions -> dozens of partitions
    ctx.set("batch_bytes", 100_000)          # several batches per call
    ctx.set("log_bytes", 8 << 20)            # several flushes per call
    with tg.KmerCounter(ctx, k, is_ds=canonical, expected_keys=2000) as kc:   # tiny hint: growth by estimate
        kc.add_records(recs)
        gk, gc = kc.dump()
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, oc)
        assert kc.size() == len(ok)
        assert kc.geometry()[1] > 1
        np.testing.assert_array_equal(kc.histo(), oracle.jf_histo(oc))
        kc.add_records(recs)                 # second pass over a table that is now large enough
        gk, gc = kc.dump()
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, 2 * oc)


@pytest.mark.parametrize("path", ["host", "device", "device_fine_overflow"])
def test_logged_count_more_partitions_than_log_bins_is_refined(ctx, oracle, data, path):
    """A table with more partitions than the 512 log bins phase 1 is fast with (on one GPU: tables beyond ~8 GB; here
    4 KiB partitions): phase 1 fills 512 COARSE bins and the replay first splits them into one segment per partition
    (k_log_refine).  A repeat k-mer that overflows its fine segment is counted directly.  Bit-exact either way."""
    _, reads = data
    reads = list(reads) + [b"ACACACACACACACACACACACACACACACACACACACACACACACACAGACACACACACACACACACACACACACACACACACACACACACACACACACAC"] * 900
    if path == "device_fine_overflow":
        # 40 MB of N: no k-mers, but the log is laid out for the input size, so the coarse bin now HOLDS the repeat
        # (no direct insert in phase 1) and it is its fine segment (1/11 of the coarse bin) that overflows in the refine
        reads = reads + [b"N" * 1_000_000] * 40
    recs, offs = tg.records_from_sequences(reads)
    k = 25
    ok, oc = oracle.jf_count(recs, k, True, 1)
    assert oc.max() > 20000                                  # the repeat: tens of thousands of occurrences of two k-mers
    ctx.set("count_mode", "log")
    # 512-slot partitions -> more than a thousand of them (at this toy size a partition's fill fluctuates far more than a
    # 16-MB partition's; the table is laid out at half the usual load to keep the smallest partitions from filling up)
    ctx.set("part_bytes", 8 << 10)
    ctx.set("kernel_timing", 1)
    ctx.kernel_times()
    with tg.KmerCounter(ctx, k, is_ds=True, expected_keys=2 * len(ok)) as kc:
        nparts = kc.geometry()[1]
        assert nparts > 512 and nparts % 512 == 0
        if path == "host":
            kc.add_records(recs)
        else:
            d = _dev_records(ctx, recs)
            kc.add_records_dev(d, recs.nbytes)
        gk, gc = kc.dump()
        kt = ctx.kernel_times()
        assert "k_log_refine" in kt and "k_log_replay" in kt and "k_log_tiles" in kt
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, oc)
        assert kc.size() == len(ok)
        if path != "host":
            kc.add_records_dev(d, recs.nbytes)               # buffers reused, counts double
            gk, gc = kc.dump()
            np.testing.assert_array_equal(gc, 2 * oc)
            ctx.dev_free(d)
    ctx.set("kernel_timing", 0)
    ctx.set("part_bytes", 16 << 20)
    ctx.set("count_mode", "auto")


@pytest.mark.parametrize("prefetch,fold", [(1, 0), (0, 0), (1, 1)])
def test_logged_count_device_path_and_stats(ctx, oracle, data, prefetch, fold):
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, 25, True, 1)
    ctx.set("count_mode", "log")
    ctx.set("part_bytes", 128 << 10)
    ctx.set("log_bytes", 8 << 20)            # the record buffer is replayed in several segments
    ctx.set("replay_fold", fold)          # duplicates of a chunk folded in shared memory before the table
    ctx.set("replay_prefetch", prefetch)
    d = _dev_records(ctx, recs)
    with tg.KmerCounter(ctx, 25, is_ds=True, expected_keys=len(ok)) as kc:
        kc.add_records_dev(d, recs.nbytes)
        gk, gc = kc.dump()
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, oc)
        # statistics straight after a logged count (the log must be flushed before any lookup)
        kc.clear()
        kc.add_records_dev(d, recs.nbytes)
        okc = oracle.KmerCounter(25, True)
        for kmer, c in zip(ok, oc):
            okc.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
        om, omean, osd = okc.coverage_stats(recs, offs)
        gm, gmean, gsd = kc.coverage_stats(recs, offs)
        np.testing.assert_array_equal(gm, om)
        np.testing.assert_array_equal(gmean.view(np.uint32), omean.view(np.uint32))
        np.testing.assert_array_equal(gsd.view(np.uint32), osd.view(np.uint32))
    ctx.dev_free(d)


def test_logged_count_bin_overflow_counts_directly(ctx, oracle):
    """two alternating k-mers repeated far beyond a bin's capacity: the overflow is counted straight into the table"""
    rng = np.random.default_rng(3)
    reads = [b"AC" * 60] * 4000 + [synth.ALPHA[rng.integers(0, 4, 90)].tobytes() for _ in range(500)]
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, 25, True, 1)
    ctx.set("count_mode", "log")
    ctx.set("part_bytes", 64 << 10)
    d = _dev_records(ctx, recs)
    with tg.KmerCounter(ctx, 25, is_ds=True, expected_keys=4 * len(ok)) as kc:
        kc.add_records_dev(d, recs.nbytes)
        gk, gc = kc.dump()
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, oc)
        assert gc.max() > 100_000
    ctx.dev_free(d)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_count_exchange_emulated(ctx, oracle, data, world):
    """N ranks on one GPU: per-rank phase 1, the all-to-all done by slicing, per-rank phase 2, all-gather by copies"""
    _, reads = data
    k = 25
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, k, True, 1)
    subcap, nparts, lp = tg.sharded.shard_geometry(world, len(ok) // world + 1000, part_bytes=64 << 10)
    assert nparts == world * lp and lp > 1
    shards = [tg.KmerCounter.sharded(ctx, k, True, subcap, nparts, r * lp, lp) for r in range(world)]
    ranges = [tg.sharded.record_range(offs, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == len(reads)
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
    nbytes_max = max(int(offs[r1] - offs[r0]) for r0, r1 in ranges)
    cap = tg.sharded.log_capacity(nbytes_max, nparts)
    logs, curs = [], []
    hpoly = ctx.dev_alloc(64)
    ctx.memset(hpoly, 0, 64)
    for r, (r0, r1) in enumerate(ranges):
        sub = recs[int(offs[r0]):int(offs[r1])]
        d = _dev_records(ctx, sub)
        keys = ctx.dev_alloc(nparts * cap * LOG_ENTRY)
        cur = ctx.dev_alloc(nparts * 4)
        ctx.memset(cur, 0, nparts * 4)
        shards[r].partition_dev(d, sub.nbytes, nparts, cap, keys, cur, hpoly)
        ctx.sync()
        ctx.dev_free(d)
        logs.append(keys)
        curs.append(cur)
    # (the homopolymer side channel accumulated over all ranks = the all-reduce of the real exchange)
    hp_host = ctx.d2h(hpoly, 64, np.uint64)
    assert hp_host[4] > 0 and hp_host[7] > 0          # the poly-A and poly-T reads of the fixture
    for dst in range(world):
        rkeys = ctx.dev_alloc(nparts * cap * LOG_ENTRY)       # [world][lp][cap]
        rcur = ctx.dev_alloc(nparts * 4)
        for src in range(world):
            ctx.d2d(rkeys, logs[src], lp * cap * LOG_ENTRY, dst_off=src * lp * cap * LOG_ENTRY, src_off=dst * lp * cap * LOG_ENTRY)
            ctx.d2d(rcur, curs[src], lp * 4, dst_off=src * lp * 4, src_off=dst * lp * 4)
        hp_d = ctx.dev_alloc(64)                       # every rank applies its own copy of the global tallies
        ctx.h2d(hp_d, hp_host)
        shards[dst].replay_log_dev(rkeys, rcur, hp_d, world, cap)
        ctx.sync()
        ctx.dev_free(hp_d)
        ctx.dev_free(rkeys)
        ctx.dev_free(rcur)
    assert sum(s.size() for s in shards) == len(ok)
    # every rank's dump is its part of the global dump
    parts = [s.dump() for s in shards]
    allk = np.concatenate([p[0] for p in parts])
    allc = np.concatenate([p[1] for p in parts])
    order = np.argsort(allk, kind="stable")
    np.testing.assert_array_equal(allk[order], ok)
    np.testing.assert_array_equal(allc[order], oc)
    # all-gather: concatenated shard slot arrays == the full table
    full = tg.KmerCounter.sharded(ctx, k, True, subcap, nparts, 0, nparts)
    fp, fbytes = full.slots_dev()
    off = 0
    for s in shards:
        sp, sb = s.slots_dev()
        ctx.d2d(fp, sp, sb, dst_off=off)
        off += sb
    assert off == fbytes
    full.set_distinct(len(ok))
    gk, gc = full.dump()
    np.testing.assert_array_equal(gk, ok)
    np.testing.assert_array_equal(gc, oc)
    okc = oracle.KmerCounter(k, True)
    for kmer, c in zip(ok, oc):
        okc.add_kmer(tg.packed_to_kmer(kmer, k), int(c))
    om, omean, osd = okc.coverage_stats(recs, offs)
    gm, gmean, gsd = full.coverage_stats(recs, offs)
    np.testing.assert_array_equal(gm, om)
    np.testing.assert_array_equal(gsd.view(np.uint32), osd.view(np.uint32))
    # a k-mer routed to the wrong shard is an error, not a silent miss
    with pytest.raises(tg.TrinityGpuError):
        shards[0].add_records(recs)
        shards[0].size()
    for t in shards + [full]:
        t.close()
    for p in logs + curs + [hpoly]:
        ctx.dev_free(p)


@pytest.mark.parametrize("world,fine_per_coarse", [(2, 1), (8, 1), (2, 4), (4, 16), (8, 2)])
def test_sharded_count_peer_exchange_emulated(ctx, oracle, data, world, fine_per_coarse):
    """The fused phase 1 + exchange (tg_count_partition_peers_dev) with N ranks emulated on one GPU: every "rank" runs
    phase 1 with the pointers of all N receive logs and writes segment [rank] of each; what peer memory adds on a real
    box is only where those pointers point.  Each owner then replays its log [src][lp][cap] with the cursor rows the
    ranks would exchange.  Result: bit-exact against the oracle, and identical to the collective exchange."""
    _, reads = data
    k = 25
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, k, True, 1)
    subcap, nparts, lp = tg.sharded.shard_geometry(world, len(ok) // world + 1000, part_bytes=64 << 10)
    assert nparts == world * lp and lp > 1 and lp & (lp - 1) == 0
    shards = [tg.KmerCounter.sharded(ctx, k, True, subcap, nparts, r * lp, lp) for r in range(world)]
    ranges = [tg.sharded.record_range(offs, r, world) for r in range(world)]
    # the k-mers travel in cbins = world * c COARSE bins; fine_per_coarse > 1: the owner splits them (tg_log_refine_dev)
    part_lp = lp
    assert lp % fine_per_coarse == 0
    c = lp // fine_per_coarse
    fine_nparts, nparts, lp = nparts, world * c, c
    bound = max(int(offs[r1] - offs[r0]) for r0, r1 in ranges)
    cap = tg.sharded.log_capacity(bound, nparts)
    rlogs = [ctx.dev_alloc(nparts * cap * LOG_ENTRY) for _ in range(world)]       # rank r's receive log [world][c][cap]
    for p in rlogs:
        ctx.memset(p, 0xEE, nparts * cap * LOG_ENTRY)                             # stale bytes must never be replayed
    curs = []
    hpoly = ctx.dev_alloc(64)
    ctx.memset(hpoly, 0, 64)
    for r, (r0, r1) in enumerate(ranges):
        sub = recs[int(offs[r0]):int(offs[r1])]
        d = _dev_records(ctx, sub)
        cur = ctx.dev_alloc(nparts * 4)
        ctx.memset(cur, 0, nparts * 4)
        shards[r].partition_peers_dev(d, sub.nbytes, nparts, cap, r, rlogs, cur, hpoly)
        ctx.sync()
        ctx.dev_free(d)
        curs.append(cur)
    hp_host = ctx.d2h(hpoly, 64, np.uint64)
    for dst in range(world):
        rcur = ctx.dev_alloc(nparts * 4)                                   # [world][lp]: row src = src's cursors of my bins
        for src in range(world):
            ctx.d2d(rcur, curs[src], lp * 4, dst_off=src * lp * 4, src_off=dst * lp * 4)
        hp_d = ctx.dev_alloc(64)
        ctx.h2d(hp_d, hp_host)
        if fine_per_coarse == 1:
            shards[dst].replay_log_dev(rlogs[dst], rcur, hp_d, world, cap)
        else:
            fcap = tg.sharded.log_capacity(world * bound, part_lp, slack=1.3)
            fkeys = ctx.dev_alloc(part_lp * fcap * LOG_ENTRY)
            fcur = ctx.dev_alloc(part_lp * 4)
            ctx.memset(fkeys, 0xEE, part_lp * fcap * LOG_ENTRY)
            ctx.memset(fcur, 0, part_lp * 4)
            _lib.check(_lib.lib().tg_log_refine_dev(ctx._h, rlogs[dst], rcur, world, c, cap, fkeys, fcur, part_lp, fcap,
                                                    dst * part_lp, fine_nparts))
            ctx.sync()
            # nothing lost, nothing invented: the fine cursors add up to the coarse ones
            assert int(ctx.d2h(fcur, part_lp * 4, np.uint32).sum()) == int(ctx.d2h(rcur, nparts * 4, np.uint32).sum())
            shards[dst].replay_log_dev(fkeys, fcur, hp_d, 1, fcap)
            ctx.sync()
            ctx.dev_free(fkeys)
            ctx.dev_free(fcur)
        ctx.sync()
        ctx.dev_free(hp_d)
        ctx.dev_free(rcur)
    assert sum(s.size() for s in shards) == len(ok)
    parts = [s.dump() for s in shards]
    allk = np.concatenate([p[0] for p in parts])
    allc = np.concatenate([p[1] for p in parts])
    order = np.argsort(allk, kind="stable")
    np.testing.assert_array_equal(allk[order], ok)
    np.testing.assert_array_equal(allc[order], oc)
    # argument checks: bins per rank must be a power of two, at most 8 ranks, rank inside
    cur = curs[0]
    with pytest.raises(tg.TrinityGpuError):
        shards[0].partition_peers_dev(rlogs[0], 0, 3 * world, cap, 0, rlogs, cur, hpoly)
    with pytest.raises(tg.TrinityGpuError):
        shards[0].partition_peers_dev(rlogs[0], 0, nparts, cap, world, rlogs, cur, hpoly)
    # a coarse log refined by the WRONG owner: every key is foreign, reported at the next sync
    if fine_per_coarse > 1 and world > 1:
        fkeys = ctx.dev_alloc(part_lp * 64 * LOG_ENTRY)
        fcur = ctx.dev_alloc(part_lp * 4)
        ctx.memset(fcur, 0, part_lp * 4)
        rcur = ctx.dev_alloc(nparts * 4)
        for src in range(world):
            ctx.d2d(rcur, curs[src], lp * 4, dst_off=src * lp * 4, src_off=0)
        _lib.check(_lib.lib().tg_log_refine_dev(ctx._h, rlogs[0], rcur, world, c, cap, fkeys, fcur, part_lp, 64,
                                                1 * part_lp, fine_nparts))
        with pytest.raises(tg.TrinityGpuError):
            ctx.sync()
        ctx.sync()
        for p in (fkeys, fcur, rcur):
            ctx.dev_free(p)
    for t in shards:
        t.close()
    for p in rlogs + curs + [hpoly]:
        ctx.dev_free(p)


def test_ipc_handle_roundtrip_api(ctx):
    """tg_ipc_export gives a 64-byte handle for a tg_dev_alloc allocation; opening it in the SAME process is refused by
    CUDA (handles are for other processes) and must surface as an error, not a crash"""
    from trinityrnaseq_b200 import _lib
    L = _lib.lib()
    p = ctx.dev_alloc(1 << 20)
    h = (C.c_uint8 * _lib.TG_IPC_HANDLE_BYTES)()
    _lib.check(L.tg_ipc_export(ctx._h, p, h))
    assert any(bytes(h))
    q = C.c_void_p()
    rc = L.tg_ipc_open(ctx._h, h, C.byref(q))
    if rc == 0:                      # some drivers allow it; then it must close cleanly
        _lib.check(L.tg_ipc_close(ctx._h, q))
    else:
        assert rc == _lib.TG_ERR_CUDA
    ctx.sync()
    ctx.dev_free(p)


def test_partition_log_overflow_is_reported(ctx, data):
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    d = _dev_records(ctx, recs)
    nbins, cap = 8, 64                                  # far too small
    keys = ctx.dev_alloc(nbins * cap * LOG_ENTRY)
    cur = ctx.dev_alloc(nbins * 4)
    hp = ctx.dev_alloc(64)
    ctx.memset(cur, 0, nbins * 4)
    ctx.memset(hp, 0, 64)
    with tg.KmerCounter(ctx, 25) as kc:
        kc.partition_dev(d, recs.nbytes, nbins, cap, keys, cur, hp)
        with pytest.raises(tg.TrinityGpuError) as e:
            ctx.sync()
        assert "overflow" in str(e.value)
    ctx.sync()                                          # the flag is cleared once reported
    for p in (d, keys, cur, hp):
        ctx.dev_free(p)


@pytest.mark.parametrize("canonical", [True, False])
def test_compacted_min2_table_gives_identical_stats(ctx, oracle, data, canonical):
    """`jellyfish dump -L 2` kept on the device: the table without its count-1 k-mers answers every coverage
    query exactly like the full table (fastaToKmerCoverageStats clamps counts below 1 to 1)"""
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, 25, canonical, 1)
    ctx.set("part_bytes", 256 << 10)
    with tg.KmerCounter(ctx, 25, is_ds=canonical, expected_keys=len(ok)) as kc:
        kc.add_records(recs)
        assert kc.count_min(1) == len(ok)
        assert kc.count_min(2) == int((oc >= 2).sum())
        assert 0 < kc.count_min(2) < len(ok)
        full = kc.coverage_stats(recs, offs, capture_coverage_info=True)
        with kc.compacted(2) as q:
            assert q.size() == int((oc >= 2).sum())
            k2, c2 = q.dump()
            np.testing.assert_array_equal(k2, ok[oc >= 2])
            np.testing.assert_array_equal(c2, oc[oc >= 2])
            part = q.coverage_stats(recs, offs, capture_coverage_info=True)
            for a, b in zip(full, part):
                np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
            # refill the same destination: same content again
            kc.compact_into(2, q)
            assert q.size() == int((oc >= 2).sum())
            # after more counting every k-mer has count >= 2: the old destination is too small and says so
            kc.add_records(recs)
            kc.compact_into(2, q)
            with pytest.raises(tg.TrinityGpuError) as e:
                q.size()
            assert "overflow" in str(e.value)
        with kc.compacted(2) as q2:
            assert q2.size() == len(ok)
        with kc.compacted(3) as q3:                    # -L 3 is NOT statistics-preserving: count-2 k-mers vanish
            assert q3.size() == int((2 * oc >= 3).sum())


def test_crowded_buckets_and_long_walks(ctx, oracle):
    """A table under pressure: every single-base variant of a few sequences, on a table tight enough (load ~0.65) that
    buckets overflow and walks are long; absent double variants are looked up too.  Counts, dumps and statistics must stay
    exact through the direct and the logged count path."""
    rng = np.random.default_rng(99)
    base = [synth.ALPHA[rng.integers(0, 4, 120)].tobytes() for _ in range(6)]
    reads = []
    for b in base:
        for rep in range(3):
            reads.append(b)
        for i in range(len(b)):
            for alt in b"ACGT":
                if alt != b[i]:
                    reads.append(b[:i] + bytes([alt]) + b[i + 1:])
    probes = list(reads)
    for b in base:                               # absent double variants: looked up, never counted
        for i in range(0, len(b) - 3, 7):
            v = bytearray(b)
            v[i] = ord("A") if v[i] != ord("A") else ord("C")
            v[i + 3] = ord("G") if v[i + 3] != ord("G") else ord("T")
            probes.append(bytes(v))
    recs, offs = tg.records_from_sequences(reads)
    precs, poffs = tg.records_from_sequences(probes)
    for mode in ("direct", "log"):
        for canonical in (True, False):
            ok, oc = oracle.jf_count(recs, 25, canonical, 1)
            ctx.set("count_mode", mode)
            ctx.set("part_bytes", 64 << 10)
            with tg.KmerCounter(ctx, 25, is_ds=canonical, expected_keys=int(len(ok) * 0.7)) as kc:     # load ~0.65
                kc.add_records(recs)
                gk, gc = kc.dump()
                np.testing.assert_array_equal(gk, ok)
                np.testing.assert_array_equal(gc, oc)
                okc = oracle.KmerCounter(25, canonical)
                for kmer, c in zip(ok, oc):
                    okc.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
                om, omean, osd, oper = okc.coverage_stats(precs, poffs, capture=True)
                gm, gmean, gsd, gper = kc.coverage_stats(precs, poffs, capture_coverage_info=True)
                np.testing.assert_array_equal(gper, oper)
                np.testing.assert_array_equal(gm, om)
                np.testing.assert_array_equal(gsd.view(np.uint32), osd.view(np.uint32))
    ctx.set("count_mode", "auto")
