"""GPU parity: every C-ABI entry point against the CPU oracle on the same seeded inputs (bit-exact)."""
import numpy as np
import pytest

import trinityrnaseq_b200 as tg
import synthdata as synth

pytestmark = pytest.mark.gpu


def _f32_bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(20251017)
    txs = synth.transcriptome(rng, 60, mean_len=900, min_len=200, max_len=4000)
    reads = synth.reads_from(rng, txs, 6000, 100, lower_rate=0.1, var_len=True)
    reads += [b"", b"ACGT", b"A" * 25, b"A" * 26, b"ACGTN" * 30, b"N" * 60, txs[0][:25], txs[0][:26],
              txs[1][:300], txs[2][:1000], txs[3][:2500], b"acgtnACGTN" * 13]
    return txs, reads


@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("k", [25, 21, 31])
def test_count_dump_histo(gpu_ctx, oracle, data, canonical, k):
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, k, canonical, 1)
    with tg.KmerCounter(gpu_ctx, k, is_ds=canonical, expected_keys=1000) as kc:   # tiny hint: forces growth
        kc.add_records(recs)
        gk, gc = kc.dump(min_count=1)
        assert kc.size() == len(ok)
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, oc)
        gk2, gc2 = kc.dump(min_count=2)
        ok2, oc2 = oracle.jf_count(recs, k, canonical, 2)
        np.testing.assert_array_equal(gk2, ok2)
        np.testing.assert_array_equal(gc2, oc2)
        np.testing.assert_array_equal(kc.histo(), oracle.jf_histo(oc))
        # counting the same buffer again doubles every count
        kc.add_records(recs)
        gk3, gc3 = kc.dump()
        np.testing.assert_array_equal(gk3, ok)
        np.testing.assert_array_equal(gc3, 2 * oc)


@pytest.mark.parametrize("ds", [True, False])
def test_cov_stats_from_reads(gpu_ctx, oracle, data, ds):
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    okc = oracle.KmerCounter(25, ds)
    okc.add_records(recs, offs)        # S3: reads of exactly k bases are skipped by the reference loader
    om, omean, osd, oper = okc.coverage_stats(recs, offs, capture=True)
    keep = [r for r in reads if len(r) != 25]
    krecs, _ = tg.records_from_sequences(keep)
    with tg.KmerCounter(gpu_ctx, 25, is_ds=ds) as kc:
        kc.add_records(krecs)
        assert kc.size() == okc.size()
        gm, gmean, gsd, gper = kc.coverage_stats(recs, offs, capture_coverage_info=True)
    np.testing.assert_array_equal(gm, om)
    np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(omean))
    np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(osd))
    np.testing.assert_array_equal(gper, oper)


def test_cov_stats_from_dump_pairs(gpu_ctx, oracle, data):
    """--kmers path: (k-mer, count) pairs incl. both strands of one k-mer, huge counts (u32 wrap) and count 0"""
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    keys, cnts = oracle.jf_count(recs, 25, False, 1)         # non-canonical dump: both strands present
    rng = np.random.default_rng(5)
    cnts = cnts.copy()
    cnts[rng.integers(0, len(cnts), 50)] = np.uint32(4000000000)
    cnts[rng.integers(0, len(cnts), 50)] = 0
    okc = oracle.KmerCounter(25, True)
    for kmer, c in zip(keys, cnts):
        okc.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
    om, omean, osd = okc.coverage_stats(recs, offs)
    with tg.KmerCounter(gpu_ctx, 25, is_ds=True) as kc:
        kc.add_kmers(keys, cnts)
        assert kc.size() == okc.size()
        gm, gmean, gsd = kc.coverage_stats(recs, offs)
    np.testing.assert_array_equal(gm, om)
    np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(omean))
    np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(osd))


@pytest.mark.parametrize("strand", [False, True])
def test_assign(gpu_ctx, oracle, data, strand):
    txs, reads = data
    rng = np.random.default_rng(11)
    names, bundles = synth.bundles_from(rng, txs)
    bundles[3] = bundles[3].lower()                    # bundles are upper-cased by the reference loader
    brecs, boffs = tg.records_from_sequences(bundles)
    # chimeric reads: halves from two different transcripts -> mixed labels, ties and the last-label quirk
    chim = [txs[i][:50] + txs[i + 1][:50] for i in range(0, 40, 2)] + [txs[i + 1][:38] + txs[i][:38] for i in range(0, 40, 2)]
    rr = reads + chim + [synth.revcomp(r) for r in chim]
    recs, offs = tg.records_from_sequences(rr)
    ot = oracle.BundleTable(25)
    ot.label(brecs, boffs)
    ob, op, osc = ot.assign(recs, offs, strand=strand)
    with tg.BundleKmerTable(gpu_ctx, 25, expected_keys=1000) as bt:
        bt.label_bundles(brecs, boffs)
        assert bt.size() == ot.size()
        gb, gp, gsc = bt.assign_reads(recs, offs, strand=strand)
    np.testing.assert_array_equal(gb, ob)
    np.testing.assert_array_equal(gsc, osc)
    assigned = ob >= 0
    np.testing.assert_array_equal(gp[assigned], op[assigned])
    assert assigned.sum() > 1000


@pytest.mark.parametrize("strand", [False, True])
@pytest.mark.parametrize("k", [25, 24])
def test_assign_both_orientations(gpu_ctx, oracle, data, strand, k):
    """The label table keeps one label per orientation of a canonical key: bundles that hold a sequence AND its
    reverse complement (in different bundles, so the two orientations carry different labels), k-mers shared by
    several bundles (highest index wins per orientation), and -- for even k -- palindromic k-mers, whose forward and
    reverse-complement look-ups are the same table entry."""
    txs, _ = data
    rng = np.random.default_rng(5)
    half = synth.ALPHA[rng.integers(0, 4, k // 2)].tobytes()
    pal = half + synth.revcomp(half) if k % 2 == 0 else half + b"A" + synth.revcomp(half)
    a, b, c = txs[0][:400], txs[1][:400], txs[2][:300]
    bundles = [a + b"X" + b[:200],                    # 0
               synth.revcomp(a)[:250] + b"X" + c,     # 1: the far end of a, reverse-complemented
               b"G" * 5 + pal + b"C" * 7 + b"X" + b,  # 2: palindrome (even k) + all of b (re-labels bundle 0's part)
               synth.revcomp(c) + b"X" + synth.revcomp(b)[:180],   # 3
               pal + b"T" + a[100:180]]               # 4: the palindrome again, and a piece of a
    brecs, boffs = tg.records_from_sequences(bundles)
    srcs = [a, b, c, synth.revcomp(a), synth.revcomp(b), synth.revcomp(c), b"G" * 5 + pal + b"C" * 7 + b[:60],
            synth.revcomp(b"G" * 5 + pal + b"C" * 7 + b[:60])]
    reads = []
    for s in srcs:
        for off in range(0, max(1, len(s) - 60), 13):
            reads.append(s[off:off + 60 + (off % 37)])
    reads += [pal, pal + b"A", synth.revcomp(pal + b"A")]
    recs, offs = tg.records_from_sequences(reads)
    ot = oracle.BundleTable(k)
    ot.label(brecs, boffs)
    ob, op, osc = ot.assign(recs, offs, strand=strand)
    with tg.BundleKmerTable(gpu_ctx, k, expected_keys=100) as bt:       # small hint: the table also has to grow
        bt.label_bundles(brecs, boffs)
        assert bt.size() == ot.size()
        gb, gp, gsc = bt.assign_reads(recs, offs, strand=strand)
    np.testing.assert_array_equal(gb, ob)
    np.testing.assert_array_equal(gsc, osc)
    assigned = ob >= 0
    np.testing.assert_array_equal(gp[assigned], op[assigned])
    assert assigned.sum() > 50 and len(set(ob[assigned].tolist())) >= 4


def test_device_resident_entry_points_with_long_reads(gpu_ctx, oracle, data):
    """tg_cov_stats_dev / tg_assign_reads_dev never synchronise with the host: the CTA-per-read kernel for reads beyond
    the warp path is launched unconditionally and finds the long reads itself.  Same results as the host-buffer entry
    points (which are checked against the oracle above), and a read that outgrows the scratch budget is an error at
    the next sync, never a wrong answer."""
    txs, reads = data
    rng = np.random.default_rng(99)
    giant = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 150_000))
    reads = list(reads) + [txs[3][:2500] * 3, giant[:40_000]]
    recs, offs = tg.records_from_sequences(reads)
    n = len(reads)
    ctx = gpu_ctx
    d_recs = ctx.dev_records_alloc(recs.nbytes)
    ctx.h2d(d_recs, recs)
    d_offs = ctx.dev_alloc(offs.nbytes)
    ctx.h2d(d_offs, offs)
    d1, d2, d3 = ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n)
    with tg.KmerCounter(ctx, 25, is_ds=True) as kc:
        kc.add_records(recs)
        hm, hmean, hsd = kc.coverage_stats(recs, offs)
        for _ in range(2):                                   # twice: the long-read list is reset per call
            kc.coverage_stats_dev(d_recs, d_offs, n, d1, d2, d3)
        ctx.sync()
        np.testing.assert_array_equal(ctx.d2h(d1, 4 * n, np.uint32), hm)
        np.testing.assert_array_equal(ctx.d2h(d2, 4 * n, np.uint32), _f32_bits(hmean))
        np.testing.assert_array_equal(ctx.d2h(d3, 4 * n, np.uint32), _f32_bits(hsd))
        om, omean, osd = _oracle_stats(oracle, recs, offs)
        np.testing.assert_array_equal(hm, om)
        np.testing.assert_array_equal(_f32_bits(hsd), _f32_bits(osd))
    names, bundles = synth.bundles_from(rng, txs)
    brecs, boffs = tg.records_from_sequences(bundles)
    with tg.BundleKmerTable(ctx, 25) as bt:
        bt.label_bundles(brecs, boffs)
        hb, hp, _ = bt.assign_reads(recs, offs, strand=False)
        d_lut = ctx.dev_alloc(bt.entropy_ok.nbytes)
        ctx.h2d(d_lut, bt.entropy_ok)
        bt.assign_reads_dev(d_recs, d_offs, n, d_lut, d1, d2, strand=False)
        ctx.sync()
        np.testing.assert_array_equal(ctx.d2h(d1, 4 * n, np.int32), hb)
        np.testing.assert_array_equal(ctx.d2h(d2, 4 * n, np.int32)[hb >= 0], hp[hb >= 0])
        ctx.dev_free(d_lut)
    # a 150 kb read against a 1 MiB scratch budget
    recs2, offs2 = tg.records_from_sequences([giant, reads[0]])
    with tg.Context(ctx.device) as small:
        small.set("long_scratch_mb", 1)
        d_r3 = small.dev_records_alloc(recs2.nbytes)
        small.h2d(d_r3, recs2)
        d_o3 = small.dev_alloc(offs2.nbytes)
        small.h2d(d_o3, offs2)
        e1, e2, e3 = small.dev_alloc(8), small.dev_alloc(8), small.dev_alloc(8)
        with tg.KmerCounter(small, 25, is_ds=True) as kc:
            kc.add_records(recs2)
            kc.coverage_stats_dev(d_r3, d_o3, 2, e1, e2, e3)
            with pytest.raises(tg.TrinityGpuError) as e:
                small.sync()
            assert "too long" in str(e.value)
            small.sync()
            # the host-buffer entry point sizes its scratch from the data and handles the same read
            gm, _, _ = kc.coverage_stats(recs2, offs2)
            assert gm[0] >= 1
    for p in (d_recs, d_offs, d1, d2, d3):
        ctx.dev_free(p)


def _oracle_stats(oracle, recs, offs):
    okc = oracle.KmerCounter(25, True)
    okc.add_records(recs, offs)
    # KmerCounter.add_records follows the reference loader (reads of exactly k bases skipped); the GPU table above counted
    # them, so add them back for an apples-to-apples comparison
    for i in range(len(offs) - 1):
        seq = bytes(recs[int(offs[i]):int(offs[i + 1]) - 1])
        if len(seq) == 25 and all(ch in b"ACGTacgt" for ch in seq):
            okc.add_kmer(seq.decode().upper(), 1)
    return okc.coverage_stats(recs, offs)


@pytest.mark.parametrize("k", [5, 7, 8, 9])
def test_short_kmers(gpu_ctx, oracle, k):
    """very short k-mers (nearly every window repeats, reads shorter than k, windows of one chunk): count, per-window coverage
    and statistics equal the oracle's."""
    rng = np.random.default_rng(k)
    txs = synth.transcriptome(rng, 5, mean_len=300, min_len=100, max_len=600)
    reads = synth.reads_from(rng, txs, 400, 60, var_len=True) + [b"A" * 40, b"ACGTN" * 9, b"", b"AC"]
    recs, offs = tg.records_from_sequences(reads)
    for canonical in (True, False):
        ok, oc = oracle.jf_count(recs, k, canonical, 1)
        with tg.KmerCounter(gpu_ctx, k, is_ds=canonical, expected_keys=len(ok) + 16) as kc:
            kc.add_records(recs)
            gk, gc = kc.dump()
            np.testing.assert_array_equal(gk, ok)
            np.testing.assert_array_equal(gc, oc)
            okc = oracle.KmerCounter(k, canonical)
            for kmer, c in zip(ok, oc):
                okc.add_kmer(tg.packed_to_kmer(kmer, k), int(c))
            om, omean, osd, oper = okc.coverage_stats(recs, offs, capture=True)
            gm, gmean, gsd, gper = kc.coverage_stats(recs, offs, capture_coverage_info=True)
            np.testing.assert_array_equal(gper, oper)
            np.testing.assert_array_equal(gm, om)
            np.testing.assert_array_equal(gmean.view(np.uint32), omean.view(np.uint32))
            np.testing.assert_array_equal(gsd.view(np.uint32), osd.view(np.uint32))


def test_held_records_one_upload_same_results(gpu_ctx, oracle, data):
    """tg_records_hold: count + statistics + assignment on ONE device copy of the reads give exactly what the separate
    uploads give (and what the oracle gives), in either order of first use, with a long read falling back to the batched
    path, and after release / a new hold."""
    txs, reads = data
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, 25, True, 1)
    okc = oracle.KmerCounter(25, True)
    for kmer, c in zip(ok, oc):
        okc.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
    om, omean, osd = okc.coverage_stats(recs, offs)
    pinned, owner = gpu_ctx.pinned((recs.nbytes,), np.uint8)
    pinned[:] = recs
    for first in ("count", "stats"):
        gpu_ctx.records_hold(pinned)
        with tg.KmerCounter(gpu_ctx, 25, is_ds=True, expected_keys=len(ok)) as kc:
            if first == "stats":            # the statistics call uploads; the table is still empty: every coverage is 1
                m0, _, _ = kc.coverage_stats(pinned, offs)
                assert int(m0.max()) <= 1
            kc.add_records(pinned)
            gk, gc = kc.dump()
            np.testing.assert_array_equal(gk, ok)
            np.testing.assert_array_equal(gc, oc)
            gm, gmean, gsd = kc.coverage_stats(pinned, offs)
            np.testing.assert_array_equal(gm, om)
            np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(omean))
            np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(osd))
            # a different array with the same content takes the ordinary path: same answers
            gm2, _, gsd2 = kc.coverage_stats(recs, offs)
            np.testing.assert_array_equal(gm2, om)
            np.testing.assert_array_equal(_f32_bits(gsd2), _f32_bits(osd))
        names, bundles = synth.bundles_from(np.random.default_rng(3), txs)
        brecs, boffs = tg.records_from_sequences(bundles)
        with tg.BundleKmerTable(gpu_ctx, 25) as bt:
            bt.label_bundles(brecs, boffs)
            hb, hp, hs = bt.assign_reads(pinned, offs, strand=False)
            gpu_ctx.records_release()
            b2, p2, s2 = bt.assign_reads(pinned, offs, strand=False)
            np.testing.assert_array_equal(hb, b2)
            np.testing.assert_array_equal(hp, p2)
            np.testing.assert_array_equal(hs, s2)
    gpu_ctx.records_release()


@pytest.mark.parametrize("batch_bytes", [64 << 20, 96 << 10])
def test_locus_order_changes_nothing_but_the_order(oracle, data, batch_bytes):
    """The per-read kernels visit the reads in LOCUS order (sorted by the smallest m-mer hash of the read) once a launch is
    large enough; here the threshold is 0, so every launch is reordered: statistics, per-window coverage and assignments
    must come back read by read exactly as the oracle computes them -- empty, short, all-N, long (CTA-per-read path) and
    lower-case reads included, in one launch and split over many small batches, through the host-buffer, the held-buffer
    and the device-resident entry points."""
    txs, reads = data
    recs, offs = tg.records_from_sequences(reads)
    with tg.Context(0) as ctx:
        ctx.set("locus_min_reads", 0)
        ctx.set("batch_bytes", batch_bytes)
        ctx.set("kernel_timing", 1)
        ok, oc = oracle.jf_count(recs, 25, True, 1)
        okc = oracle.KmerCounter(25, True)
        for kmer, c in zip(ok, oc):
            okc.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
        om, omean, osd, oper = okc.coverage_stats(recs, offs, capture=True)
        names, bundles = synth.bundles_from(np.random.default_rng(5), txs)
        brecs, boffs = tg.records_from_sequences(bundles)
        ot = oracle.BundleTable(25)
        ot.label(brecs, boffs)
        ob, op, os_ = ot.assign(recs, offs, strand=False)
        with tg.KmerCounter(ctx, 25, is_ds=True, expected_keys=len(ok)) as kc, tg.BundleKmerTable(ctx, 25) as bt:
            kc.add_records(recs)
            bt.label_bundles(brecs, boffs)
            ctx.kernel_times()
            gm, gmean, gsd, gper = kc.coverage_stats(recs, offs, capture_coverage_info=True)
            assert "k_locus_tiles" in ctx.kernel_times()
            np.testing.assert_array_equal(gper, oper)
            np.testing.assert_array_equal(gm, om)
            np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(omean))
            np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(osd))
            gb, gp, gs = bt.assign_reads(recs, offs, strand=False)
            np.testing.assert_array_equal(gb, ob)
            np.testing.assert_array_equal(gs, os_)
            np.testing.assert_array_equal(gp[ob >= 0], op[ob >= 0])
            # held buffer: one device copy, one launch over all reads
            pinned, owner = ctx.pinned((recs.nbytes,), np.uint8)
            pinned[:] = recs
            ctx.records_hold(pinned)
            hm, hmean, hsd = kc.coverage_stats(pinned, offs)
            hb, hp, hs = bt.assign_reads(pinned, offs, strand=False)
            ctx.records_release()
            np.testing.assert_array_equal(hm, om)
            np.testing.assert_array_equal(_f32_bits(hsd), _f32_bits(osd))
            np.testing.assert_array_equal(hb, ob)
            np.testing.assert_array_equal(hs, os_)
            # the knob off: same answers
            ctx.set("locus_order", 0)
            ctx.kernel_times()
            nm, nmean, nsd = kc.coverage_stats(recs, offs)
            assert "k_locus_tiles" not in ctx.kernel_times()
            np.testing.assert_array_equal(nm, om)
            np.testing.assert_array_equal(_f32_bits(nsd), _f32_bits(osd))


@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("k", [25, 31, 9])
def test_count_read_by_read_in_locus_order(oracle, data, canonical, k):
    """tg_count_records_dev: reads counted one by one in locus order, straight into the table (no log): same dump and
    histogram as the oracle -- homopolymer runs, reads longer than one warp segment, N's and empty reads included --
    with the order computed per call, pinned (computed once, reused), and switched off."""
    _, reads = data
    reads = list(reads) + [b"A" * 400, b"T" * 90 + b"G" * 90, b"AC" * 300, b"\x00\xff"]
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, k, canonical, 1)
    with tg.Context(0) as ctx:
        ctx.set("locus_min_reads", 0)
        d_recs = ctx.dev_records_alloc(recs.nbytes); ctx.h2d(d_recs, recs)
        d_offs = ctx.dev_alloc(offs.nbytes); ctx.h2d(d_offs, offs)
        for mode in ("fresh", "pinned", "off"):
            if mode == "pinned":
                ctx.records_pin_dev(d_recs, d_offs, len(reads))
            if mode == "off":
                ctx.records_pin_dev()
                ctx.set("locus_order", 0)
            with tg.KmerCounter(ctx, k, is_ds=canonical, expected_keys=len(ok) + 64) as kc:
                for rep in range(2 if mode == "pinned" else 1):
                    kc.add_read_records_dev(d_recs, d_offs, len(reads))
                gk, gc = kc.dump()
                np.testing.assert_array_equal(gk, ok)
                np.testing.assert_array_equal(gc, oc * (2 if mode == "pinned" else 1))


@pytest.mark.parametrize("floor", [2, 3])
def test_count_floor_view_equals_rebuilt_dump_table(gpu_ctx, oracle, data, floor):
    """tg_table_set_count_floor(n): the statistics read the count table as `jellyfish dump -L n` would leave it, without
    rebuilding it.  Same medians / means / stdevs / per-window coverage as (a) the oracle's table loaded from the -L n dump and
    (b) our own materialised -L n table; counting, dump and histo do not see the floor; floor 0 restores the full table."""
    _, reads = data
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = oracle.jf_count(recs, 25, True, floor)
    okc = oracle.KmerCounter(25, True)
    for kmer, c in zip(ok, oc):
        okc.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
    om, omean, osd, oper = okc.coverage_stats(recs, offs, capture=True)
    with tg.KmerCounter(gpu_ctx, 25, is_ds=True) as kc:
        kc.add_records(recs)
        full = kc.coverage_stats(recs, offs)
        kc.set_count_floor(floor)
        gm, gmean, gsd, gper = kc.coverage_stats(recs, offs, capture_coverage_info=True)
        np.testing.assert_array_equal(gper, oper)
        np.testing.assert_array_equal(gm, om)
        np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(omean))
        np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(osd))
        q = kc.compacted(floor)
        qm, qmean, qsd = q.coverage_stats(recs, offs)
        q.close()
        np.testing.assert_array_equal(qm, om)
        np.testing.assert_array_equal(_f32_bits(qsd), _f32_bits(osd))
        ak, ac = oracle.jf_count(recs, 25, True, 1)
        gk, gc = kc.dump()                                  # the dump is not a read path of the view
        np.testing.assert_array_equal(gk, ak)
        np.testing.assert_array_equal(gc, ac)
        kc.set_count_floor(0)
        again = kc.coverage_stats(recs, offs)
        for a, b in zip(full, again):
            np.testing.assert_array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


def _k32_reads(rng):
    txs = synth.transcriptome(rng, 20, mean_len=600, min_len=200, max_len=2000)
    reads = synth.reads_from(rng, txs, 1500, 100, lower_rate=0.1, var_len=True)
    # the 32-mers a 64-bit key has no tag bit for: poly-A (the all-zero key) and its reverse complement, in runs of every
    # kind -- exactly k, k + 1, long, lower case, broken after 31 bases, behind an N -- plus the other homopolymers
    reads += [b"A" * 32, b"a" * 33, b"A" * 120, b"T" * 32, b"T" * 77, b"A" * 31, b"A" * 31 + b"C" + b"A" * 40,
              b"T" * 31 + b"G" + b"T" * 33, b"C" * 45, b"G" * 45, b"ACGT" * 4 + b"A" * 36 + b"TTTT" + b"A" * 34 + b"N" + b"A" * 32,
              b"", b"ACGT", b"N" * 40, txs[0][:32], txs[0][:33], txs[1][:400], txs[2][:1500]]
    return txs, reads


@pytest.mark.parametrize("canonical", [True, False])
def test_k32_count_dump_histo_stats(gpu_ctx, oracle, canonical):
    """k = 32 (Inchworm/src/KmerCounter.cpp:15-17 allows it): the planes fill all 64 bits of the key, poly-A lives in the
    table's zero-key slot.  Count (with growth: the zero-key slot moves with the rehash), dump, dump -L 2, histo, sum of
    counts, per-window coverage and statistics (host-buffer, held and device-resident entry points), the -L 2 view, and
    the pair loader -- all against the oracle."""
    rng = np.random.default_rng(3232)
    _, reads = _k32_reads(rng)
    recs, offs = tg.records_from_sequences(reads)
    k = 32
    ok, oc = oracle.jf_count(recs, k, canonical, 1)
    assert ok[0] == 0                                    # poly-A is there
    okc = oracle.KmerCounter(k, canonical)
    for kmer, c in zip(ok, oc):
        okc.add_kmer(tg.packed_to_kmer(kmer, k), int(c))
    om, omean, osd, oper = okc.coverage_stats(recs, offs, capture=True)
    ctx = gpu_ctx
    with tg.KmerCounter(ctx, k, is_ds=canonical, expected_keys=500) as kc:       # tiny hint: forces growth
        kc.add_records(recs)
        assert kc.size() == len(ok)
        gk, gc = kc.dump()
        np.testing.assert_array_equal(gk, ok)
        np.testing.assert_array_equal(gc, oc)
        ok2, oc2 = oracle.jf_count(recs, k, canonical, 2)
        gk2, gc2 = kc.dump(min_count=2)
        np.testing.assert_array_equal(gk2, ok2)
        np.testing.assert_array_equal(gc2, oc2)
        np.testing.assert_array_equal(kc.histo(), oracle.jf_histo(oc))
        assert kc.count_sum() == int(oc.astype(np.uint64).sum())
        gm, gmean, gsd, gper = kc.coverage_stats(recs, offs, capture_coverage_info=True)
        np.testing.assert_array_equal(gper, oper)
        np.testing.assert_array_equal(gm, om)
        np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(omean))
        np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(osd))
        # held buffer and device-resident entry points
        pinned, owner = ctx.pinned((recs.nbytes,), np.uint8)
        pinned[:] = recs
        ctx.records_hold(pinned)
        hm, hmean, hsd = kc.coverage_stats(pinned, offs)
        ctx.records_release()
        np.testing.assert_array_equal(hm, om)
        np.testing.assert_array_equal(_f32_bits(hsd), _f32_bits(osd))
        n = len(reads)
        d_recs = ctx.dev_records_alloc(recs.nbytes)
        ctx.h2d(d_recs, recs)
        d_offs = ctx.dev_alloc(offs.nbytes)
        ctx.h2d(d_offs, offs)
        d1, d2, d3 = ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n)
        kc.coverage_stats_dev(d_recs, d_offs, n, d1, d2, d3)
        ctx.sync()
        np.testing.assert_array_equal(ctx.d2h(d1, 4 * n, np.uint32), om)
        np.testing.assert_array_equal(ctx.d2h(d2, 4 * n, np.uint32), _f32_bits(omean))
        np.testing.assert_array_equal(ctx.d2h(d3, 4 * n, np.uint32), _f32_bits(osd))
        # the `dump -L 2` view and its materialised twin
        okc2 = oracle.KmerCounter(k, canonical)
        for kmer, c in zip(ok2, oc2):
            okc2.add_kmer(tg.packed_to_kmer(kmer, k), int(c))
        om2, _, osd2 = okc2.coverage_stats(recs, offs)
        kc.set_count_floor(2)
        fm, _, fsd = kc.coverage_stats(recs, offs)
        np.testing.assert_array_equal(fm, om2)
        np.testing.assert_array_equal(_f32_bits(fsd), _f32_bits(osd2))
        q = kc.compacted(2)
        qk, qc = q.dump()
        qm, _, qsd = q.coverage_stats(recs, offs)
        q.close()
        np.testing.assert_array_equal(qk, ok2)
        np.testing.assert_array_equal(qc, oc2)
        np.testing.assert_array_equal(qm, om2)
        np.testing.assert_array_equal(_f32_bits(qsd), _f32_bits(osd2))
        kc.set_count_floor(0)
        # counting the same buffer again, from the device copy, doubles every count
        kc.add_records_dev(d_recs, recs.nbytes)
        gk3, gc3 = kc.dump()
        np.testing.assert_array_equal(gk3, ok)
        np.testing.assert_array_equal(gc3, 2 * oc)
        for p in (d_recs, d_offs, d1, d2, d3):
            ctx.dev_free(p)
    # the pair loader (--kmers): a non-canonical dump loaded into a table of either kind, poly-A and poly-T both present
    fk, fc = oracle.jf_count(recs, k, False, 1)
    assert fk[0] == 0 and fk[-1] == np.uint64(0xFFFFFFFFFFFFFFFF)
    okl = oracle.KmerCounter(k, canonical)
    for kmer, c in zip(fk, fc):
        okl.add_kmer(tg.packed_to_kmer(kmer, k), int(c))
    lm, lmean, lsd = okl.coverage_stats(recs, offs)
    with tg.KmerCounter(ctx, k, is_ds=canonical, expected_keys=500) as kc:
        kc.add_kmers(fk, fc)
        assert kc.size() == okl.size()
        gm, gmean, gsd = kc.coverage_stats(recs, offs)
    np.testing.assert_array_equal(gm, lm)
    np.testing.assert_array_equal(_f32_bits(gmean), _f32_bits(lmean))
    np.testing.assert_array_equal(_f32_bits(gsd), _f32_bits(lsd))


def test_k32_is_refused_where_the_key_tag_is_needed(gpu_ctx):
    """label tables, sharded tables and the partitioned / read-by-read count entry points keep k <= 31; k = 33 nowhere"""
    with pytest.raises(tg.TrinityGpuError):
        tg.BundleKmerTable(gpu_ctx, 32)
    with pytest.raises(tg.TrinityGpuError):
        tg.KmerCounter.sharded(gpu_ctx, 32, True, 1024, 8, 0, 4)
    with pytest.raises(tg.TrinityGpuError):
        tg.KmerCounter(gpu_ctx, 33)
