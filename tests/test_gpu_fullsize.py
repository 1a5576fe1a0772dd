"""GPU, BASELINE.json configs[1] at FULL size (10 M pairs = 20 M reads x 100 bp, generated on the device): the oracle
cannot finish this in seconds, so the checks are size-independent properties of the path.

  * conservation: the counts add up to the number of valid 25-mer windows of the input, counted independently on the
    host with numpy from the raw bytes (no k-mer code involved);
  * two independent kernels, one answer: the direct-insert path and the partitioned (log + replay) path give the same
    distinct count and the same count histogram (a checksum of checksums over 150 M keys);
  * linearity: counting the same reads twice doubles every count (histogram shifted c -> 2c);
  * idempotence / equivalence of the query side: statistics against the `dump -L 2` table are bit-identical to those
    against the full table, every median and mean is >= 1 (counts are clamped to 1), and the n = 76 windows of a
    100-bp read bound mean by the largest count;
  * read -> bundle assignment at the same size: every assigned read has pct >= the threshold semantics allow, the
    double-stranded run assigns at least the reads the strand-specific run assigns.
"""
import numpy as np
import pytest

import trinityrnaseq_b200 as tg
from trinityrnaseq_b200 import _lib

pytestmark = pytest.mark.gpu

K = 25
PAIRS = 10_000_000
READ_LEN = 100


def _valid_windows(recs, k):
    """windows of k consecutive ACGTacgt bytes, by cumulative sums over chunks (pure numpy)"""
    lut = np.zeros(256, dtype=np.uint8)
    for ch in b"ACGTacgt":
        lut[ch] = 1
    total, n, step = 0, len(recs), 64 << 20
    for a in range(0, n, step):
        b = min(n, a + step + k - 1)
        bad = (lut[recs[a:b]] == 0).astype(np.int32)
        cs = np.concatenate(([0], np.cumsum(bad, dtype=np.int32)))
        w = cs[k:] - cs[:-k]                       # invalid bytes inside the window starting at a + i
        lim = min(step, n - a - k + 1)             # windows that START in this chunk
        if lim > 0:
            total += int((w[:lim] == 0).sum())
    return total


@pytest.fixture(scope="module")
def full(gpu_ctx):
    import bench
    ctx = gpu_ctx
    tx, tx_offs, tx_cum = bench.make_transcriptome(20_000, bench.SEED)
    d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, PAIRS, READ_LEN, seed=bench.SEED)
    nreads = 2 * PAIRS
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(READ_LEN + 1)
    d_offs = ctx.dev_alloc(offs.nbytes)
    ctx.h2d(d_offs, offs)
    yield {"ctx": ctx, "tx": tx, "tx_offs": tx_offs, "d_recs": d_recs, "nbytes": nbytes, "nreads": nreads, "d_offs": d_offs}
    ctx.dev_free(d_recs)
    ctx.dev_free(d_offs)


def test_count_conservation_two_paths_and_linearity(full):
    ctx, d_recs, nbytes = full["ctx"], full["d_recs"], full["nbytes"]
    recs = ctx.d2h(d_recs, nbytes, np.uint8)
    windows = _valid_windows(recs, K)
    assert 0.9 * 2 * PAIRS * (READ_LEN - K + 1) < windows <= 2 * PAIRS * (READ_LEN - K + 1)
    expected = int(full["tx_offs"][-1]) + int(2 * PAIRS * READ_LEN * 0.005 * K * 0.68) + (1 << 20)
    weights = np.arange(_lib.TG_HISTO_BINS, dtype=np.uint64)

    ctx.set("count_mode", "log")
    with tg.KmerCounter(ctx, K, True, expected_keys=expected) as kc:
        kc.add_records_dev(d_recs, nbytes)
        h_log, n_log = kc.histo(), kc.size()
        # the last bin holds every count beyond 10000 (jellyfish histo): those few k-mers are dumped and added exactly
        _, big = kc.dump(min_count=10001, sorted_=False)
        assert len(big) == int(h_log[-1])
        assert int((h_log[:-1] * weights[:-1]).sum()) + int(big.astype(np.uint64).sum()) == windows      # conservation
        assert int(h_log.sum()) == n_log
        kc.add_records_dev(d_recs, nbytes)                        # linearity
        h2 = kc.histo()
        assert kc.size() == n_log
        top = (_lib.TG_HISTO_BINS - 2) // 2
        np.testing.assert_array_equal(h2[2:2 * top + 1:2], h_log[1:top + 1])
        assert int(h2[1:2 * top:2].sum()) == 0
    ctx.set("count_mode", "direct")
    with tg.KmerCounter(ctx, K, True, expected_keys=expected) as kc:
        kc.add_records_dev(d_recs, nbytes)
        np.testing.assert_array_equal(kc.histo(), h_log)          # two kernels, one answer
        assert kc.size() == n_log
    ctx.set("count_mode", "auto")


def test_stats_full_vs_dump_L2_table_and_bounds(full):
    ctx, d_recs, nbytes, nreads, d_offs = full["ctx"], full["d_recs"], full["nbytes"], full["nreads"], full["d_offs"]
    expected = int(full["tx_offs"][-1]) + int(2 * PAIRS * READ_LEN * 0.005 * K * 0.68) + (1 << 20)
    d = [ctx.dev_alloc(4 * nreads) for _ in range(3)]
    with tg.KmerCounter(ctx, K, True, expected_keys=expected) as kc:
        kc.add_records_dev(d_recs, nbytes)
        kc.coverage_stats_dev(d_recs, d_offs, nreads, *d)
        ctx.sync()
        med = ctx.d2h(d[0], 4 * nreads, np.uint32)
        mean = ctx.d2h(d[1], 4 * nreads, np.float32)
        sd_bits = ctx.d2h(d[2], 4 * nreads, np.uint32)
        assert med.min() >= 1 and mean.min() >= 1.0 and np.isfinite(mean).all()
        _, big = kc.dump(min_count=10001, sorted_=False)
        top = int(big.max()) if len(big) else int(np.nonzero(kc.histo())[0].max())
        assert med.max() <= top and mean.max() <= top
        # a read's windows were all counted from the same reads: at least one occurrence each unless broken by N
        q = kc.compacted(2, load=0.40)
        q.coverage_stats_dev(d_recs, d_offs, nreads, *d)
        ctx.sync()
        np.testing.assert_array_equal(ctx.d2h(d[0], 4 * nreads, np.uint32), med)
        np.testing.assert_array_equal(ctx.d2h(d[1], 4 * nreads, np.float32).view(np.uint32), mean.view(np.uint32))
        np.testing.assert_array_equal(ctx.d2h(d[2], 4 * nreads, np.uint32), sd_bits)
        q.close()
    for p in d:
        ctx.dev_free(p)


def test_assignment_full_size_properties(full):
    import bench
    ctx, d_recs, nreads, d_offs = full["ctx"], full["d_recs"], full["nreads"], full["d_offs"]
    brecs, boffs, ncontigs = bench.make_bundles(full["tx"], full["tx_offs"], bench.SEED + 1)
    nb = len(boffs) - 1
    d_b = ctx.dev_records_alloc(brecs.nbytes)
    ctx.h2d(d_b, brecs)
    d_bo = ctx.dev_alloc(boffs.nbytes)
    ctx.h2d(d_bo, boffs)
    d_best, d_pct = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
    with tg.BundleKmerTable(ctx, K, expected_keys=int(full["tx_offs"][-1]) + (1 << 20)) as bt:
        d_lut = ctx.dev_alloc(bt.entropy_ok.nbytes)
        ctx.h2d(d_lut, bt.entropy_ok)
        bt.label_bundles_dev(d_b, brecs.nbytes, d_bo, nb)
        out = {}
        for strand in (True, False):
            bt.assign_reads_dev(d_recs, d_offs, nreads, d_lut, d_best, d_pct, strand=strand)
            ctx.sync()
            out[strand] = (ctx.d2h(d_best, 4 * nreads, np.int32), ctx.d2h(d_pct, 4 * nreads, np.int32))
        bt.assign_reads_dev(d_recs, d_offs, nreads, d_lut, d_best, d_pct, strand=False)     # idempotent
        ctx.sync()
        np.testing.assert_array_equal(ctx.d2h(d_best, 4 * nreads, np.int32), out[False][0])
        ctx.dev_free(d_lut)
    for strand, (best, pct) in out.items():
        assert best.min() >= -1 and best.max() < nb
        a = best >= 0
        assert pct[a].min() >= 1                                   # score >= 1 of 76 windows rounds to >= 1 %
        assert pct[a].max() <= (100 if strand else 200)            # DS can exceed 100
    # reads come from both strands of the transcripts: the forward-only run assigns about half of what DS assigns
    n_ss, n_ds = int((out[True][0] >= 0).sum()), int((out[False][0] >= 0).sum())
    assert n_ds > 0.95 * nreads and 0.35 * n_ds < n_ss <= n_ds
    # a read assigned by its forward k-mers alone keeps a bundle with at least that support in the DS run
    both = (out[True][0] >= 0) & (out[False][0] >= 0)
    assert (out[False][1][both] >= out[True][1][both]).all()
    for p in (d_b, d_bo, d_best, d_pct):
        ctx.dev_free(p)
