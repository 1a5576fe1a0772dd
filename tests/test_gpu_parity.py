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
