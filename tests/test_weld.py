"""GraphFromFasta weldmer counting (Chrysalis/analysis/GraphFromFasta.cc:1412-1424,
NonRedKmerTable.cc:162-200): the oracle restatement against the counts of the unmodified reference classes
(tests/golden/weld_counts.json, made by tests/golden/make_golden_weld.py), the GPU path against both."""
import json
import os

import numpy as np
import pytest

import weldcase

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "weld_counts.json")))


def _records(reads):
    import trinityrnaseq_b200 as tg
    return tg.records_from_sequences(reads)


@pytest.mark.parametrize("kk", [48, 33, 40])
def test_oracle_weld_counts_equal_reference_classes(oracle, kk):
    reads, cands = weldcase.weld_case(seed=11, kk=kk)
    recs, offs = _records(reads)
    got = oracle.weld_count(cands, kk, recs, offs)
    np.testing.assert_array_equal(got, np.asarray(GOLD[str(kk)], np.int32))
    assert got[-1] == 0                                   # a candidate with an N matches no read window
    assert got[-5] == got[3] == got[-4]                   # duplicate candidates share one counter


@pytest.mark.gpu
@pytest.mark.parametrize("kk", [48, 33, 40])
def test_gpu_weld_counts(gpu_ctx, oracle, kk):
    import trinityrnaseq_b200 as tg
    reads, cands = weldcase.weld_case(seed=11, kk=kk)
    recs, offs = _records(reads)
    want = np.asarray(GOLD[str(kk)], np.int32)
    with tg.WeldmerTable(gpu_ctx, cands, kk) as wt:
        wt.add_records(recs)
        np.testing.assert_array_equal(wt.counts(), want)
        d = gpu_ctx.dev_records_alloc(recs.nbytes)
        gpu_ctx.h2d(d, recs)
        wt.add_records_dev(d, recs.nbytes)                # the same reads again, from HBM: every counter doubles
        gpu_ctx.sync()
        np.testing.assert_array_equal(wt.counts(), 2 * want)
        gpu_ctx.dev_free(d)
    # a second, larger random case straight against the oracle (no golden): many candidates, small batches
    reads2, cands2 = weldcase.weld_case(seed=5, kk=kk, nreads=6000)
    recs2, offs2 = _records(reads2)
    gpu_ctx.set("batch_bytes", 96 << 10)
    try:
        with tg.WeldmerTable(gpu_ctx, cands2, kk) as wt:
            wt.add_records(recs2)
            np.testing.assert_array_equal(wt.counts(), oracle.weld_count(cands2, kk, recs2, offs2))
    finally:
        gpu_ctx.set("batch_bytes", 64 << 20)


@pytest.mark.gpu
def test_gpu_weld_rejects_bad_arguments(gpu_ctx):
    import trinityrnaseq_b200 as tg
    with pytest.raises(Exception):
        tg.WeldmerTable(gpu_ctx, [b"A" * 20], 20)         # kk outside 33..48
    with tg.WeldmerTable(gpu_ctx, [], 48) as wt:          # no candidates: nothing to count, nothing to crash
        wt.add_records(np.frombuffer(b"ACGT" * 30 + b"\n", dtype=np.uint8))
        assert len(wt.counts()) == 0
