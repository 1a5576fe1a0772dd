"""torchrun worker of tests/test_gpu_multi.py: ShardedKmerCounter over REAL NCCL / NVLink peer memory, one process per GPU,
checked against the CPU oracle on the concatenated reads.  argv: exchange (peer|collective) fold (0|1) coarse_bins nreads"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist
    import synthdata as synth
    import trinityrnaseq_b200 as tg
    from trinityrnaseq_b200 import sharded
    from oracle import oracle_py as orc

    exchange, fold, coarse, nreads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = 25
    rng = np.random.default_rng(4242)
    txs = synth.transcriptome(rng, 60, mean_len=700, min_len=200, max_len=2500)
    reads = synth.reads_from(rng, txs, nreads, 100, var_len=True)
    # hot spots and edge cases: a poly-A tail on every 5th read (homopolymer side channel), one k-mer carried by ~3 % of
    # the reads (adapter-dimer style skew), dinucleotide repeats, empty and too-short reads, N runs
    adapter = b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"
    for i in range(0, len(reads), 5):
        reads[i] = reads[i][:60] + b"A" * 40
    for i in range(3, len(reads), 33):
        reads[i] = reads[i][:50] + adapter
    reads += [b"AC" * 60, b"", b"ACGT" * 5, b"ACGTN" * 20, txs[0][:25], b"T" * 300]
    recs, offs = tg.records_from_sequences(reads)
    ok, oc = orc.jf_count(recs, k, True, 1)
    r0, r1 = sharded.record_range(offs, rank, world)
    mine = recs[int(offs[r0]):int(offs[r1])]
    sub_offs = offs[r0:r1 + 1] - offs[r0]
    ctx = tg.Context(local)
    d = ctx.dev_records_alloc(max(mine.nbytes, 1))
    if mine.nbytes:
        ctx.h2d(d, mine)
    eng = sharded.DeviceEngine(ctx, k, True)
    sc = sharded.ShardedKmerCounter(eng, expected_keys_per_rank=len(ok) // world + 4096, part_bytes=64 << 10, exchange=exchange,
                                    max_exchange_bins=(world * coarse) if coarse else sharded.MAX_EXCHANGE_BINS,
                                    replay_fold=bool(fold))
    assert sc.exchange == exchange, (sc.exchange, exchange)
    for rep in (1, 2):              # the second batch reuses every buffer: counts double
        sc.add_records_dev(d, mine.nbytes)
        assert sc.exchange == exchange, "fell back from the requested exchange"
        assert sc.size() == len(ok)
        np.testing.assert_array_equal(sc.histo(), orc.jf_histo(rep * oc))
        lk, lc = sc.dump_local()
        parts = [None] * world
        dist.all_gather_object(parts, (lk, lc))
        ks, cs = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        order = np.argsort(ks, kind="stable")
        np.testing.assert_array_equal(ks[order], ok)
        np.testing.assert_array_equal(cs[order], rep * oc)
    okc = orc.KmerCounter(k, True)
    for key, c in zip(ok.tolist(), oc.tolist()):
        okc.add_kmer(tg.packed_to_kmer(key, k), 2 * c)
    om, omean, osd = okc.coverage_stats(recs, offs)
    for min_count in (1, 2, 3):
        full = sc.replicate(min_count=min_count)
        fk, fc = full.dump()
        keep = (2 * oc) >= min_count
        np.testing.assert_array_equal(fk, ok[keep])
        np.testing.assert_array_equal(fc, (2 * oc)[keep])
        if r1 > r0 and min_count <= 2:
            gm, gmean, gsd = full.coverage_stats(mine, sub_offs)
            np.testing.assert_array_equal(gm, om[r0:r1])
            np.testing.assert_array_equal(gmean.view(np.uint32), omean[r0:r1].view(np.uint32))
            np.testing.assert_array_equal(gsd.view(np.uint32), osd[r0:r1].view(np.uint32))
    # routed lookups: no replica, keys to the owners and counts back -- the same statistics, also through the -L 2 floor
    if mine.nbytes:
        d_offs = ctx.dev_alloc(sub_offs.nbytes)
        ctx.h2d(d_offs, np.ascontiguousarray(sub_offs, dtype=np.uint64))
    else:
        d_offs = ctx.dev_alloc(8)
    n_mine = r1 - r0
    d_m, d_a, d_s = ctx.dev_alloc(4 * max(n_mine, 1)), ctx.dev_alloc(4 * max(n_mine, 1)), ctx.dev_alloc(4 * max(n_mine, 1))
    okc2 = orc.KmerCounter(k, True)                      # what a table rebuilt from `dump -L 3` answers
    for key, c in zip(ok.tolist(), oc.tolist()):
        if 2 * c >= 3:
            okc2.add_kmer(tg.packed_to_kmer(key, k), 2 * c)
    om3, omean3, osd3 = okc2.coverage_stats(recs, offs)
    for min_count, (wm, wmean, wsd) in ((1, (om, omean, osd)), (3, (om3, omean3, osd3))):
        sc.coverage_stats_routed_dev(d, mine.nbytes, d_offs, n_mine, d_m, d_a, d_s, min_count=min_count)
        if n_mine:
            np.testing.assert_array_equal(ctx.d2h(d_m, 4 * n_mine, np.uint32), wm[r0:r1])
            np.testing.assert_array_equal(ctx.d2h(d_a, 4 * n_mine, np.uint32), wmean[r0:r1].view(np.uint32))
            np.testing.assert_array_equal(ctx.d2h(d_s, 4 * n_mine, np.uint32), wsd[r0:r1].view(np.uint32))
    sc.close()
    dist.barrier()
    if rank == 0:
        print(f"MP_SHARDED_OK world={world} exchange={exchange} fold={fold} coarse={coarse} distinct={len(ok)}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
