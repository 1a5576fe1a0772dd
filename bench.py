#!/usr/bin/env python
"""bench.py -- Trinity k-mer hot path on B200: 25-mers/s counted + queried.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the hot path over the whole synthetic read set: clear the table, count every canonical
25-mer (jellyfish count stage), keep the k-mers seen at least twice (`jellyfish dump -L 2`, on the device), then per-read
k-mer coverage statistics for every read against that table (fastaToKmerCoverageStats).  Workload at N=1 = BASELINE.json configs[1]: 10 M PE 2x100 bp reads from a random
20 k-transcript set (20 M reads, 2.0 Gbase; 1.52 G k-mer positions counted + 1.52 G queried per step).

  value     device-resident throughput: reads already in HBM, CUDA events on the library's stream
  e2e       the same step through the host-buffer C ABI (tg_count_reads + tg_cov_stats on pinned host
            arrays): H2D of reads/offsets and D2H of the per-read results inside the timed region
  roofline  every kernel of one step with its CUDA-event time, algorithmic HBM bytes and fraction of the measured
            HBM peak (the top one is reported as `roofline`); random_access = count / lookup rates against the
            GUPS probes measured in this run
  cpu_baseline  the reference's own CPU tool (oracle/_ref/fastaToKmerCoverageStats, else the C oracle) on a
            bounded sample of the same reads, timed on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch on the default workload, from the committed `ncu --set full`
# captures (profiles/README.md); filled in when a capture of the current kernels exists
# dram__bytes_read.sum + dram__bytes_write.sum per launch, configs[1] on one B200, from the round-2 `ncu --set full` captures
# (profiles/r02_ncu_full_count_locus_statsmin2_raw.csv, profiles/r02_ncu_full_stats_raw.csv)
NCU_TRAFFIC_BYTES = {"k_log_tiles": 2.077e9 + 11.844e9, "k_log_replay": 28.805e9 + 5.265e9, "k_cov_stats": 45.089e9 + 2.441e9,
                     "k_locus_tiles": 2.262e9 + 0.162e9, "locus_sort": 0.0, "k_cov_stats_long": 0.0}

K = 25
METRIC = "25-mers/sec counted+queried"
UNIT = "kmers/s"
SEED = 20251017


# ---------------------------------------------------------------------------------------------------------
# synthetic transcriptome (host, numpy) -- reads themselves are generated on the device from it
# ---------------------------------------------------------------------------------------------------------
def make_transcriptome(ntx, seed, sigma=2.0):
    rng = np.random.default_rng(seed)
    lens = np.clip(np.round(rng.lognormal(np.log(1500), 0.6, ntx)), 300, 10000).astype(np.int64)
    offs = np.zeros(ntx + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    tx = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(offs[-1]))]
    w = rng.lognormal(0.0, sigma, ntx) * lens          # expression x length = share of fragments
    cum = np.cumsum(w / w.sum())
    cum_u64 = np.minimum(cum * 2.0 ** 64, 2.0 ** 64 - 2048).astype(np.uint64)
    cum_u64[-1] = np.uint64(2 ** 64 - 1)
    return tx, offs, cum_u64


def make_bundles(tx, tx_offs, seed, max_per_bundle=25):
    """Inchworm-bundle records cut from the transcriptome (SURVEY §8d, C4 shape): contigs = consecutive pieces of
    every transcript with length ~ lognormal(ln 500, 0.7) clipped to [100, 20000]; bundles of 1..25 consecutive
    contigs joined by 'X' (Chrysalis/analysis/CreateIwormFastaBundle.cc:56-68).  -> (record buffer, offs, ncontigs)"""
    rng = np.random.default_rng(seed)
    ntx = len(tx_offs) - 1
    pieces = []
    for t in range(ntx):
        a, e = int(tx_offs[t]), int(tx_offs[t + 1])
        while a < e:
            n = int(np.clip(round(rng.lognormal(np.log(500), 0.7)), 100, 20000))
            if e - a - n < 100:
                n = e - a
            pieces.append((a, a + n))
            a += n
    out, offs, i = [], [0], 0
    xs = np.frombuffer(b"X", dtype=np.uint8)
    nl = np.frombuffer(b"\n", dtype=np.uint8)
    while i < len(pieces):
        m = int(rng.integers(1, max_per_bundle + 1))
        grp = pieces[i:i + m]
        i += m
        n = 0
        for j, (a, e) in enumerate(grp):
            if j:
                out.append(xs); n += 1
            out.append(tx[a:e]); n += e - a
        out.append(nl); n += 1
        offs.append(offs[-1] + n)
    return np.concatenate(out), np.asarray(offs, dtype=np.uint64), len(pieces)


def r2t_cpu_reference(tx_small, offs_small, cum_small, read_len, nreads, seed, threads):
    """oracle/_ref/ReadsToTranscripts on a small instance of the same shape (its k-mer table build -- a sort of
    25-byte strings -- is part of the tool, so the instance is bounded on both sides).  -> reads/s or None"""
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "ReadsToTranscripts")
    if not os.path.exists(ref_bin):
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synthdata
    rng = np.random.default_rng(seed)
    txs = [tx_small[int(offs_small[i]):int(offs_small[i + 1])].tobytes() for i in range(len(offs_small) - 1)]
    reads = synthdata.reads_from(rng, txs, nreads, read_len, err=0.005, n_rate=0.001)
    brecs, boffs, _ = make_bundles(tx_small, offs_small, seed)
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "reads.fa"), "wb") as f:
            for i, r in enumerate(reads):
                f.write(b">r%d/1\n%s\n" % (i, r))
        with open(os.path.join(td, "bundles.fa"), "wb") as f:
            for i in range(len(boffs) - 1):
                f.write(b">s_%d 10\n" % i)
                f.write(brecs[int(boffs[i]):int(boffs[i + 1])].tobytes())
        t0 = time.perf_counter()
        subprocess.run([ref_bin, "-i", os.path.join(td, "reads.fa"), "-f", os.path.join(td, "bundles.fa"), "-o",
                        os.path.join(td, "out"), "-t", str(threads), "-max_mem_reads", "50000000", "-p", "10"],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True,
                       env={**os.environ, "OMP_NUM_THREADS": str(threads)})
        dt = time.perf_counter() - t0
    return nreads / dt, dt


def bench_r2t(ctx, tg, tx, tx_offs, d_recs, nbytes, d_offs, offs_host, recs_host, nreads, read_len, steps, want_cpu,
              world=1, max_over_ranks=lambda x: x):
    """ReadsToTranscripts on the same reads (BASELINE metric, second half): label the bundle k-mers, assign every
    read.  Device-resident and host-buffer timings; the reads are the step's 20 M reads, the bundles are cut from
    the same transcriptome."""
    from trinityrnaseq_b200 import _lib
    L = _lib.lib()
    brecs, boffs, ncontigs = make_bundles(tx, tx_offs, SEED + 1)
    nb = len(boffs) - 1
    d_b = ctx.dev_records_alloc(brecs.nbytes)
    ctx.h2d(d_b, brecs)
    d_bo = ctx.dev_alloc(boffs.nbytes)
    ctx.h2d(d_bo, boffs)
    d_best, d_pct = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
    bt = tg.BundleKmerTable(ctx, K, expected_keys=int(tx_offs[-1]) + (1 << 20))
    d_lut = ctx.dev_alloc(bt.entropy_ok.nbytes)
    ctx.h2d(d_lut, bt.entropy_ok)

    def dev_step():
        bt.clear()
        bt.label_bundles_dev(d_b, brecs.nbytes, d_bo, nb)
        bt.assign_reads_dev(d_recs, d_offs, nreads, d_lut, d_best, d_pct, strand=False)

    dev_step()
    ctx.sync()
    ctx.set("kernel_timing", 1)
    ctx.kernel_times()
    ctx.timer_start()
    for _ in range(steps):
        dev_step()
    ms = max_over_ranks(ctx.timer_stop() / steps)
    kt = ctx.kernel_times()
    ctx.set("kernel_timing", 0)
    best_d = ctx.d2h(d_best, 4 * nreads, np.int32)
    pct_d = ctx.d2h(d_pct, 4 * nreads, np.int32)
    labelled = bt.size()
    # host-buffer C ABI (H2D of bundles + reads, D2H of the assignments inside the timed region)
    best_h, o1 = ctx.pinned((nreads,), np.int32)
    pct_h, o2 = ctx.pinned((nreads,), np.int32)
    brecs_p, o3 = ctx.pinned((brecs.nbytes,), np.uint8)
    brecs_p[:] = brecs

    def host_step():
        bt.clear()
        _lib.check(L.tg_label_bundles(bt._h, brecs_p.ctypes.data, boffs.ctypes.data, nb, 0))
        _lib.check(L.tg_assign_reads(bt._h, recs_host.ctypes.data, offs_host.ctypes.data, nreads, 0,
                                     bt.entropy_ok.ctypes.data, best_h.ctypes.data, pct_h.ctypes.data, None))
    host_step()
    t0 = time.perf_counter()
    host_step()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(best_d, best_h) and np.array_equal(pct_d[best_d >= 0], pct_h[best_h >= 0]), "r2t dev/host mismatch"
    nwin = read_len - K + 1
    lookups = 2 * nreads * nwin          # forward + reverse-complement pass (no -strand)
    out = {"metric": "reads/sec ReadsToTranscripts", "workload": f"configs[3] shape on this step's reads: {nreads} reads x "
           f"{read_len} bp against {ncontigs} contigs in {nb} bundles cut from the same transcriptome, double-stranded, -p 10"
           + (f"; per GPU, {world} GPUs, reads sharded by rank, the 1.2 GB label table built on every GPU (replicas only: "
              f"the path has no exchange step)" if world > 1 else ""),
           "value": world * nreads / (ms / 1e3), "unit": "reads/s", "ms_per_step": ms, "n_gpus": world,
           "e2e": {"value": world * nreads / e2e_s, "unit": "reads/s", "ms_per_step": e2e_s * 1e3,
                   "h2d_bytes_per_step": int(brecs.nbytes + boffs.nbytes + nbytes + offs_host.nbytes),
                   "d2h_bytes_per_step": int(8 * nreads)},
           "bundle_kmers": int(labelled), "reads_assigned": int((best_d >= 0).sum()),
           "kernels": {k_: {"ms": round(v[0] / steps, 3)} for k_, v in kt.items()},
           "lookups_per_s": lookups / (kt.get("k_assign", (ms * steps, 0))[0] / steps / 1e3)}
    if want_cpu:
        ntx_s = 1500
        try:
            r = r2t_cpu_reference(tx[:int(tx_offs[ntx_s])], tx_offs[:ntx_s + 1], None, read_len, 200_000, SEED + 2,
                                  os.cpu_count() or 1)
        except (OSError, subprocess.SubprocessError):
            r = None
        if r:
            out["cpu_baseline"] = {"value": round(r[0], 1), "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference",
                                   "seconds": round(r[1], 2),
                                   "sample": f"oracle/_ref/ReadsToTranscripts -t {os.cpu_count()} on 200000 reads x {read_len} bp "
                                             f"against the bundles of the first {ntx_s} transcripts (table build included)"}
    for p_ in (d_b, d_bo, d_best, d_pct, d_lut):
        ctx.dev_free(p_)
    bt.close()
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------------
# CPU reference leg
# ---------------------------------------------------------------------------------------------------------
def write_sample_fasta(path, recs, read_len, nreads):
    """`>r<8 digits>/1` + sequence, one line each, straight from the record buffer (vectorised: 20 M records in seconds)"""
    stride = read_len + 1
    with open(path, "wb") as f:
        for a in range(0, nreads, 2_000_000):
            n = min(2_000_000, nreads - a)
            rows = np.empty((n, 13 + stride), dtype=np.uint8)
            rows[:, 0] = ord(">"); rows[:, 1] = ord("r")
            idx = np.arange(a, a + n, dtype=np.int64)
            for d in range(8):
                rows[:, 9 - d] = (idx // 10 ** d % 10 + ord("0")).astype(np.uint8)
            rows[:, 10] = ord("/"); rows[:, 11] = ord("1"); rows[:, 12] = ord("\n")
            rows[:, 13:] = recs[a * stride:(a + n) * stride].reshape(n, stride)
            rows.tofile(f)


def cpu_reference_run(sample_recs, read_len, nreads, threads=6):
    """Time the reference CPU tool on `nreads` reads: count (--kmers_from_reads) + stats in one process, exactly the
    two stages of a step.  Returns (kmers_per_s, seconds, kind, cores)."""
    positions = 2 * nreads * (read_len - K + 1)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "fastaToKmerCoverageStats")
    if os.path.exists(ref_bin):
        try:
            with tempfile.TemporaryDirectory() as td:
                fa = os.path.join(td, "sample.fa")
                write_sample_fasta(fa, sample_recs, read_len, nreads)
                t0 = time.perf_counter()
                with open(os.path.join(td, "out.stats"), "wb") as out:
                    subprocess.run([ref_bin, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", str(K), "--num_threads",
                                    str(threads), "--DS"], stdout=out, stderr=subprocess.DEVNULL, check=True)
                dt = time.perf_counter() - t0
            return positions / dt, dt, "reference", threads    # the tool caps itself at 6 threads (MAX_THREADS)
        except (OSError, subprocess.SubprocessError) as e:     # e.g. a binary built for another libc: use the port
            sys.stderr.write(f"bench: reference binary failed ({e}); timing the oracle port instead\n")
    from oracle import oracle_py as orc
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(read_len + 1)
    buf = sample_recs[: nreads * (read_len + 1)]
    t0 = time.perf_counter()
    kc = orc.KmerCounter(K, True)
    kc.add_records(buf, offs)
    kc.coverage_stats(buf, offs)
    dt = time.perf_counter() - t0
    return positions / dt, dt, "port", 1


def _stats_cols(out_bytes):
    """what downstream reads of a stats file: columns acc / median / mean / stdev, sorted"""
    rows = [l.split(b"\t")[:4] for l in out_bytes.split(b"\n") if l and not l.startswith(b"acc\t")]
    return sorted(rows)


def cli_end_to_end(sample_recs, read_len, n_small, n_large):
    """The drop-in executable, whole process, file in / file out (SURVEY §8d "end-to-end rate"): (1) on the CPU-baseline
    sample our tool and the reference tool must print the same statistics (columns 1-4, sorted: the reference's line
    order and tid depend on its threads); (2) wall clock of our tool on a larger file.  CUDA context creation, FASTA
    parsing and text formatting are all inside."""
    ours = os.path.join(ROOT, "trinityrnaseq_b200", "bin", "fastaToKmerCoverageStats")
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "fastaToKmerCoverageStats")
    if not os.path.exists(ours):
        return None
    out = {}
    with tempfile.TemporaryDirectory() as td:
        def run(tool, fa, threads=None):
            cmd = [tool, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", str(K), "--DS"]
            if threads:
                cmd += ["--num_threads", str(threads)]
            t0 = time.perf_counter()
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
            return r.stdout, time.perf_counter() - t0
        fa = os.path.join(td, "small.fa")
        write_sample_fasta(fa, sample_recs, read_len, n_small)
        mine, t_mine = run(ours, fa)
        out["sample_reads"] = n_small
        out["ours_seconds"] = round(t_mine, 3)
        if os.path.exists(ref_bin):
            theirs, t_ref = run(ref_bin, fa, 6)
            out["reference_seconds"] = round(t_ref, 3)
            out["identical_cols_1_4_sorted"] = _stats_cols(mine) == _stats_cols(theirs)
        if n_large > n_small:
            fa2 = os.path.join(td, "large.fa")
            write_sample_fasta(fa2, sample_recs, read_len, n_large)
            # stdout goes to a file, as in the pipeline (`... > left.fa.K25.stats`, util/insilico_read_normalization.pl:846)
            stats_path = os.path.join(td, "large.stats")
            # TRINITY_GPU_TRACE=1: the tool's own phase clock on stderr -- which part of the wall time is CUDA context creation
            # and process teardown (neither is the tool's to shorten; both vary from box to box) and which is its own work
            t0 = time.perf_counter()
            w0 = time.time()
            with open(stats_path, "wb") as so:
                r = subprocess.run([ours, "--reads", fa2, "--kmers_from_reads", fa2, "--kmer_size", str(K), "--DS"], stdout=so,
                                   stderr=subprocess.PIPE, check=True, env=dict(os.environ, TRINITY_GPU_TRACE="1"))
            t_big = time.perf_counter() - t0
            w1 = time.time()
            positions = 2 * n_large * (read_len - K + 1)
            out["large"] = {"reads": n_large, "seconds": round(t_big, 3), "value": round(positions / t_big, 1), "unit": UNIT,
                            "fasta_bytes": os.path.getsize(fa2), "output_bytes": os.path.getsize(stats_path)}
            phases = {}
            for line in r.stderr.decode(errors="replace").splitlines():
                if not line.startswith("[trace]"):
                    continue
                f = line.split(None, 3)
                try:
                    if f[1] == "wall":
                        phases["wall_" + f[3].strip()] = float(f[2])
                    elif f[2] == "s":
                        phases[f[3].strip()] = float(f[1])
                except (IndexError, ValueError):
                    pass
            if "wall_main" in phases and "wall_exit" in phases:
                out["large"]["phases_s"] = {
                    "exec_to_main": round(phases["wall_main"] - w0, 3),
                    "cuda_context_open": phases.get("device context(s) open"),
                    "table_counted_at": phases.get("k-mer table loaded / counted"),
                    "output_complete_at": phases.get("output complete"),
                    "exit_to_reaped": round(w1 - phases["wall_exit"], 3)}
    return out


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per GPU (configs[1]: 10 M)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--ntx", type=int, default=20_000)
    ap.add_argument("--cpu-sample-reads", type=int, default=300_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cli-reads", type=int, default=20_000_000, help="reads in the file the drop-in executable is timed on")
    ap.add_argument("--no-gups", action="store_true")
    ap.add_argument("--no-r2t", action="store_true", help="skip the ReadsToTranscripts measurement")
    ap.add_argument("--count-mode", default="auto", choices=["auto", "direct", "log"])
    ap.add_argument("--stats-table", default="auto", choices=["auto", "min2", "full"],
                    help="min2: statistics read the device-side `jellyfish dump -L 2` table -- what the normalisation pipeline "
                         "feeds fastaToKmerCoverageStats (util/insilico_read_normalization.pl:45,641); its build is part of the "
                         "timed step; full: the count table itself (bit-identical statistics, see DESIGN.md); auto = min2 at "
                         "every GPU count")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "collective"],
                    help="multi-GPU k-mer exchange: peer = phase 1 stores into the owners' logs over NVLink (fused), "
                         "collective = NCCL all-to-all of the bins; auto = peer when peer memory maps")
    ap.add_argument("--replay-fold", type=int, default=-1, help="fold duplicate k-mers per replay chunk (-1 = auto: 4+ GPUs)")
    ap.add_argument("--exchange-bins", type=int, default=0, help="coarse bins the k-mers are exchanged in (0 = default)")
    ap.add_argument("--batches", type=int, default=1,
                    help="several GPUs: a rank's reads are counted (and, with --stats routed, queried) in this many slices, so that "
                         "the exchange logs of one slice fit beside a large shard (BASELINE configs[4]: 125 M reads per GPU)")
    ap.add_argument("--stats", default="replica", choices=["replica", "routed"],
                    help="several GPUs: statistics on an all-gathered replica of the -L 2 shards (default) or through routed "
                         "lookups against the sharded table (no replica: for tables too large to replicate)")
    ap.add_argument("--routed", action="store_true",
                    help="several GPUs: also time the statistics through ROUTED lookups (no replica: keys to the owners, counts "
                         "back) and check them against the replica path bit for bit")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    read_len, npairs = args.read_len, args.pairs
    nreads = 2 * npairs
    nwin = read_len - K + 1
    positions_per_step = 2 * nreads * nwin            # counted + queried, per GPU
    # the step follows the pipeline at every GPU count: count -> `dump -L 2` (kept on the device) -> statistics on that table
    min_count = 1 if args.stats_table == "full" else 2
    config = {"workload": f"configs[1]: synthetic {npairs / 1e6:g}M PE 2x{read_len} bp reads from a random "
                          f"{args.ntx}-transcript set, k=25 canonical count + fastaToKmerCoverageStats, per GPU",
              "reads_per_gpu": nreads, "k": K, "unit_definition": "k-mer window positions counted + positions queried",
              "l2_policy": "inputs (2.0 GB reads, multi-GB table) exceed the 126 MB L2; no flush needed",
              "table_sharding": (f"hash-sharded over {world} GPUs: k-mer log all-to-all (NCCL), shards all-gathered "
                                 f"for the statistics") if world > 1 else "single table",
              "stats_table": "k-mers with count >= 2 (`jellyfish dump -L 2`, as the normalisation pipeline does: "
                             "util/insilico_read_normalization.pl:45,641): a count-floor view of the count table on one GPU, "
                             "compacted + all-gathered shards on several; statistics bit-identical to a rebuilt -L 2 table "
                             "(asserted in the run)" if min_count == 2 else "full count table",
              "read_order": "per-read kernels visit the reads in locus order (signature + radix sort inside the timed step)",
              "count_mode": args.count_mode}

    if args.impl == "reference":
        if rank != 0:
            return
        return run_reference_arm(args, config, world, W)

    import trinityrnaseq_b200 as tg
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = tg.Context(local_rank)
    ctx.set("count_mode", args.count_mode)
    if args.replay_fold >= 0 and world == 1:
        ctx.set("replay_fold", args.replay_fold)
    info = ctx.info()
    tx, tx_offs, tx_cum = make_transcriptome(args.ntx, SEED)
    d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, npairs, read_len, seed=SEED + 7919 * rank)
    stride = read_len + 1
    offs_host, offs_owner = ctx.pinned((nreads + 1,), np.uint64)
    offs_host[:] = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(stride)
    d_offs = ctx.dev_alloc(offs_host.nbytes)
    ctx.h2d(d_offs, offs_host)
    d_med, d_mean, d_sd = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)

    # table sized once from the shape of the data: the transcriptome's own k-mers are shared by all ranks, error
    # k-mers (~ K novel k-mers per substitution) are private to each rank's reads
    err_keys = int(nreads * read_len * 0.005 * K * 0.68)
    expected_total = int(tx_offs[-1]) + world * err_keys + (1 << 20)
    state = {"q": None}
    if world == 1:
        kc = tg.KmerCounter(ctx, K, is_ds=True, expected_keys=expected_total)

        def count_dev(recs_ptr):
            kc.clear()
            kc.add_records_dev(recs_ptr, nbytes)

        # one GPU: `dump -L 2` is a VIEW of the count table (tg_table_set_count_floor): a k-mer below the floor reads as
        # absent, which is exactly what the statistics of the rebuilt -L 2 table are -- nothing is materialised.  (Several
        # GPUs must materialise it: the compacted shards are what is all-gathered.)
        kc.set_count_floor(min_count)

        def query_table():
            return kc
        table_info = kc.info
    else:
        from trinityrnaseq_b200 import sharded
        eng = sharded.DeviceEngine(ctx, K, True)
        sc = sharded.ShardedKmerCounter(eng, expected_keys_per_rank=expected_total // world + 1, exchange=args.exchange,
                                        max_exchange_bins=args.exchange_bins or sharded.MAX_EXCHANGE_BINS,
                                        replay_fold=None if args.replay_fold < 0 else bool(args.replay_fold))
        kc = sc.table

        import ctypes
        B = max(1, args.batches)
        # counting slices: whole tiles (the count kernels work on a flat stream; a slice that ended inside a tile would have
        # that tile's windows counted again by the next slice); query slices: whole reads
        tile = 8192
        cuts_b = [min(nbytes, (nbytes * b // B + tile - 1) // tile * tile) for b in range(B)] + [nbytes]
        cuts_r = [nreads * b // B // 16 * 16 for b in range(B)] + [nreads]     # (16 reads: slices start 16-byte aligned for TMA)

        def at(ptr, off):
            return ctypes.c_void_p(ptr.value + int(off))

        def count_dev(recs_ptr):
            sc.clear()
            for b in range(B):
                n_b = cuts_b[b + 1] - cuts_b[b]
                if n_b:
                    sc.add_records_dev(at(recs_ptr, cuts_b[b]), n_b, max_windows=(n_b // stride + 2) * nwin)

        def query_table():
            return sc.replicate(min_count=min_count, load=0.40)

        def stats_routed(recs_ptr):
            # fixed-stride records: the offsets of any slice of reads, relative to its first byte, are the first entries of d_offs
            for b in range(B):
                r0, r1 = cuts_r[b], cuts_r[b + 1]
                if r1 > r0:
                    sc.coverage_stats_routed_dev(at(recs_ptr, r0 * stride), (r1 - r0) * stride, d_offs, r1 - r0, at(d_med, 4 * r0),
                                                 at(d_mean, 4 * r0), at(d_sd, 4 * r0), min_count=min_count)

        def table_info():
            return {"capacity": sc.subcap * sc.nparts, "distinct": sc.size()}

    # The reads of a step are declared immutable (pinned) and their locus order is queued first, on the library's second
    # stream: it is computed -- every step anew -- beside the count, and the statistics wait for it on the device.
    if world > 1:
        ctx.records_pin_dev(d_recs, d_offs, nreads)

    def device_step():
        if world > 1 and args.stats != "routed":
            # several GPUs: the order is computed on the second stream, in the shadow of the exchange (one GPU: the count
            # kernels leave no room beside them -- measured: 78.3 ms with, 76.5 ms without -- so the statistics call computes it)
            ctx.locus_prepare_dev(K, recompute=True)
        count_dev(d_recs)
        if world > 1 and args.stats == "routed":
            stats_routed(d_recs)
        else:
            query_table().coverage_stats_dev(d_recs, d_offs, nreads, d_med, d_mean, d_sd)

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------------------
    for _ in range(W):
        device_step()
    tinfo = table_info()
    qinfo = query_table().info() if not (world > 1 and args.stats == "routed") else tinfo
    if world > 1:
        config["count_batches"] = max(1, args.batches)
        config["statistics"] = ("routed lookups against the sharded table (no replica)" if args.stats == "routed"
                                else "replica of the -L 2 shards, all-gathered per step")
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        device_step()
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None       # (sampled during the device-timed region only: every nvidia-smi
    launches = ctx.launch_count() - launches0             # query takes driver locks that the host-synchronous e2e path would feel)
    ms = max_over_ranks(ms)
    value = world * positions_per_step * args.steps / (ms / 1e3)

    # per-kernel device times of one more step (CUDA events around every launch, on the launching stream)
    ctx.set("kernel_timing", 1)
    ctx.kernel_times()
    device_step()
    ktimes = ctx.kernel_times()
    ctx.set("kernel_timing", 0)

    # multi-GPU: host-clock milliseconds per phase of one more step (a sync on both sides of every phase, so the
    # sum exceeds the pipelined step; it says where the time goes)
    phases = None
    if world > 1:
        sc.profile = {}
        device_step()
        barrier()
        phases = {k_: round(max_over_ranks(v), 2) for k_, v in sorted(sc.profile.items())}
        sc.profile = None
        config["table_sharding"] = config["table_sharding"].replace(
            "k-mer log all-to-all (NCCL)",
            "k-mers stored into the owners' logs over NVLink peer memory by the partition kernel itself"
            if sc.exchange == "peer" else "k-mer log all-to-all (NCCL)")
        config["exchange"] = sc.exchange

    routed = None
    if world > 1 and args.routed:
        d_m2, d_a2, d_s2 = ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads), ctx.dev_alloc(4 * nreads)
        sc.coverage_stats_routed_dev(d_recs, nbytes, d_offs, nreads, d_m2, d_a2, d_s2, min_count=min_count)     # warm-up
        barrier()
        sc.profile = {}
        t0 = time.perf_counter()
        sc.coverage_stats_routed_dev(d_recs, nbytes, d_offs, nreads, d_m2, d_a2, d_s2, min_count=min_count)
        barrier()
        rms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        rph = {k_: round(max_over_ranks(v), 2) for k_, v in sorted(sc.profile.items())}
        sc.profile = None
        same = (np.array_equal(ctx.d2h(d_m2, 4 * nreads, np.uint32), ctx.d2h(d_med, 4 * nreads, np.uint32)) and
                np.array_equal(ctx.d2h(d_s2, 4 * nreads, np.uint32), ctx.d2h(d_sd, 4 * nreads, np.uint32)))
        assert same, "routed lookups != replica lookups"
        routed = {"ms_synced_phases": rms, "phases_ms_synced": rph, "identical_to_replica_path": bool(same),
                  "nvlink_bytes_per_lookup": 12, "lookups_per_gpu": nreads * nwin}
        for p_ in (d_m2, d_a2, d_s2):
            ctx.dev_free(p_)

    # ---- parity inside the bench (at the full size, on however many GPUs) ----------------------------------------
    # (1) conservation: the sum of all counts in the (sharded) table == the number of valid 25-mer windows of all reads,
    #     counted by an independent kernel straight from the ASCII; (2) on 2 GPUs: the all-reduced histogram of the table
    #     sharded over NVLink == the histogram of ONE single-GPU table counting both ranks' reads.
    count_dev(d_recs)
    barrier()
    local_sum = kc.count_sum()
    local_valid = ctx.valid_windows_dev(d_recs, nbytes, K)
    if dist is not None:
        t2 = torch.tensor([local_sum, local_valid], device="cuda", dtype=torch.int64)
        dist.all_reduce(t2)
        total_sum, total_valid = int(t2[0].item()), int(t2[1].item())
    else:
        total_sum, total_valid = local_sum, local_valid
    assert total_sum == total_valid, f"conservation violated: {total_sum} counted vs {total_valid} valid windows"
    parity = {"sum_of_counts": total_sum, "valid_windows": total_valid}
    if world == 2:
        sharded_histo = sc.histo()
        if rank == 0:
            with tg.KmerCounter(ctx, K, is_ds=True, expected_keys=expected_total) as one:
                for r_ in range(world):
                    d_o, nb_o = ctx.synth_reads_dev(tx, tx_offs, tx_cum, npairs, read_len, seed=SEED + 7919 * r_)
                    one.add_records_dev(d_o, nb_o)
                    ctx.sync()
                    ctx.dev_free(d_o)
                single_histo = one.histo()
            assert np.array_equal(np.asarray(sharded_histo), np.asarray(single_histo)), "sharded histogram != single-GPU histogram"
            parity["histogram_2gpu_equals_1gpu"] = True
        barrier()

    # ---- end-to-end through the host-buffer C ABI ---------------------------------------------------------
    recs_host, recs_owner = ctx.pinned((nbytes,), np.uint8)
    ctx.d2h(d_recs, recs_host)
    med_h, o1 = ctx.pinned((nreads,), np.uint32)
    mean_h, o2 = ctx.pinned((nreads,), np.float32)
    sd_h, o3 = ctx.pinned((nreads,), np.float32)
    from trinityrnaseq_b200 import _lib
    L = _lib.lib()
    d_stage = ctx.dev_records_alloc(nbytes) if world > 1 else None

    def host_step():
        # one upload per step: the reads are declared immutable for the duration of the step (tg_records_hold), the count
        # uploads them (overlapped with its first kernel) and the statistics find the device copy in place
        if world == 1:
            ctx.records_hold(recs_host)
            kc.clear()
            _lib.check(L.tg_count_reads(kc._h, recs_host.ctypes.data, nbytes, 1))
            q = query_table()
            _lib.check(L.tg_cov_stats(q._h, recs_host.ctypes.data, offs_host.ctypes.data, nreads, 1, med_h.ctypes.data,
                                      mean_h.ctypes.data, sd_h.ctypes.data, None))
            ctx.records_release()
        else:
            ctx.h2d(d_stage, recs_host)          # the sharded count takes the rank's reads from HBM
            ctx.h2d(d_offs, offs_host)
            count_dev(d_stage)
            if args.stats == "routed":
                stats_routed(d_stage)
            else:
                query_table().coverage_stats_dev(d_stage, d_offs, nreads, d_med, d_mean, d_sd)
            ctx.d2h(d_med, med_h)
            ctx.d2h(d_mean, mean_h)
            ctx.d2h(d_sd, sd_h)

    e2e_steps = args.steps
    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * positions_per_step * e2e_steps / e2e_s

    # parity spot check inside the bench: device-resident and host-buffer paths must agree bit for bit
    med_d = ctx.d2h(d_med, 4 * nreads, np.uint32)
    sd_d = ctx.d2h(d_sd, 4 * nreads, np.uint32)
    assert np.array_equal(med_d, med_h) and np.array_equal(sd_d, sd_h.view(np.uint32)), "device vs host path mismatch"
    if min_count > 1 and world == 1:
        # ... and the `dump -L 2` VIEW must give exactly the statistics of a materialised -L 2 table (rebuilt on the device
        # from the count table), with the locus order switched off for the cross-check
        q2 = kc.compacted(min_count, load=0.40)
        ctx.set("locus_order", 0)
        q2.coverage_stats_dev(d_recs, d_offs, nreads, d_med, d_mean, d_sd)
        ctx.sync()
        ctx.set("locus_order", 1)
        assert np.array_equal(ctx.d2h(d_med, 4 * nreads, np.uint32), med_h), "-L 2 view != materialised -L 2 table (median)"
        assert np.array_equal(ctx.d2h(d_sd, 4 * nreads, np.uint32), sd_h.view(np.uint32)), "-L 2 view != materialised -L 2 table (stdev)"
        q2.close()

    r2t = None
    if not args.no_r2t:
        if world > 1:
            barrier()
        r2t = bench_r2t(ctx, tg, tx, tx_offs, d_recs, nbytes, d_offs, offs_host, recs_host, nreads, read_len,
                        max(1, min(args.steps, 3)), (not args.no_cpu_baseline) and rank == 0, world, max_over_ranks)

    if world > 1:
        sc.close()                       # collective: unmap the peer logs before anybody frees its own
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline ---------------------------------------------------------------------------------------------
    # Algorithmic HBM bytes are SURVEY §8(d)'s per-unit figures: counting = one 16-B slot read-modify-write
    # per k-mer occurrence = 64 B; a lookup = one 32-B sector; `dump -L 2` = the count table read once (16 B/slot) + 64 B
    # per kept k-mer.  Counting is two kernels (k_log_tiles: reads -> super-k-mer log; k_log_replay: log -> table, plus
    # k_log_refine beyond 512 partitions) that only make sense together, so the figure is attributed to the STAGE and the
    # stage's time is the sum of its kernels' CUDA-event times.  `roofline` reports the stage with the largest time.
    peak, peak_kind = load_peaks()
    count_positions = nreads * nwin
    table_bytes = tinfo["capacity"] * 16 // world
    stage_of = {"k_log_tiles": "count", "k_log_replay": "count", "k_log_refine": "count", "k_flat_tiles<COUNT>": "count",
                "k_log_plan": "count", "k_rehash": "dump_L2", "k_cov_stats": "stats", "k_cov_stats_long": "stats",
                "k_locus_tiles": "stats", "locus_sort": "stats"}
    stage_bytes = {"count": count_positions * 64 + nbytes,
                   "dump_L2": table_bytes + qinfo["distinct"] * 64 // world,
                   "stats": count_positions * 32 + nbytes + 12 * nreads}
    stage_units = {"count": count_positions, "dump_L2": tinfo["capacity"] // world, "stats": count_positions}
    kernels, stages = [], {}
    step_kernel_ms = sum(v[0] for v in ktimes.values())
    for name, (kms, n) in sorted(ktimes.items(), key=lambda kv: -kv[1][0]):
        st = stage_of.get(name, "other")
        kernels.append({"kernel": name, "stage": st, "ms": round(kms, 3), "launches": n, "share": round(kms / step_kernel_ms, 3)})
        stages.setdefault(st, {"ms": 0.0, "kernels": []})
        stages[st]["ms"] += kms
        stages[st]["kernels"].append(name)
    for st, e in stages.items():
        e["ms"] = round(e["ms"], 3)
        e["share"] = round(e["ms"] / step_kernel_ms, 3)
        if st in stage_bytes and e["ms"] > 0:
            e["units"] = int(stage_units[st])
            e["algorithmic_bytes"] = int(stage_bytes[st])
            e["achieved_gbs"] = round(stage_bytes[st] / (e["ms"] / 1e3) / 1e9, 1)
            e["frac_of_hbm_peak"] = round(e["achieved_gbs"] / peak, 4)
    top_name = max((st for st in stages if st in stage_bytes), key=lambda st: stages[st]["ms"])
    top = stages[top_name]
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the stage's kernels from the `ncu --set full`
    # capture of exactly this workload (profiles/README.md names the file); null for any other shape
    ncu_traffic = NCU_TRAFFIC_BYTES
    default_shape = (world == 1 and npairs == 10_000_000 and read_len == 100 and args.ntx == 20_000 and min_count == 2)
    traffic = None
    if default_shape and all(k_ in ncu_traffic for k_ in top["kernels"] if not k_.startswith("k_log_plan")):
        traffic = sum(ncu_traffic[k_] for k_ in top["kernels"] if not k_.startswith("k_log_plan"))
    roofline = {"bound": "hbm", "kernel": " + ".join(top["kernels"]), "stage": top_name, "achieved": top.get("achieved_gbs"),
                "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": top.get("frac_of_hbm_peak"), "traffic": traffic,
                "units_per_launch": top.get("units"), "kernel_ms": top["ms"], "stages": stages, "kernels": kernels,
                "note": "achieved = algorithmic bytes of the stage (SURVEY 8d per-unit figure x units) / sum of its kernels' "
                        "CUDA-event times in one step; traffic = ncu DRAM bytes of the same kernels, one launch each, bytes"}
    if not args.no_gups:
        slots = max(tinfo["capacity"] // world, 1 << 29)         # >= 8 GiB of 16-B slots
        nops = 1 << 30
        g = {}
        for mode, name in ((0, "load16"), (1, "load8_red"), (2, "cas_red")):
            gms = ctx.gups(slots, nops, mode, reps=2)
            g[name] = {"gops": round(nops / (gms / 1e3) / 1e9, 2), "ms": round(gms, 2)}
        ra = {"table_gib": round(slots * 16 / 2 ** 30, 1), **g}
        count_ms = stages.get("count", {}).get("ms", 0)
        stats_ms = ktimes.get("k_cov_stats", (0, 0))[0]
        if count_ms:
            ra["count_gkmers_s"] = round(count_positions / count_ms / 1e6, 2)
            ra["count_vs_dram_random_rmw"] = round(count_positions / count_ms / 1e6 / g["load8_red"]["gops"], 3)
        if stats_ms:
            ra["stats_gkmers_s"] = round(count_positions / stats_ms / 1e6, 2)
            ra["stats_vs_dram_random_load"] = round(count_positions / stats_ms / 1e6 / g["load16"]["gops"], 3)
        roofline["random_access"] = ra

    cpu = None
    if not args.no_cpu_baseline:
        ns = min(args.cpu_sample_reads, nreads)
        v, dt, kind, cores = cpu_reference_run(recs_host, read_len, ns)
        cpu = {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": kind, "seconds": round(dt, 2),
               "sample": f"first {ns} reads of the same synthetic set: fastaToKmerCoverageStats --kmers_from_reads + "
                         f"stats, --num_threads {cores} (the tool's own cap), host has {os.cpu_count()} cores"}

    cli = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cli = cli_end_to_end(recs_host, read_len, min(args.cpu_sample_reads, nreads), min(args.cli_reads, nreads))
        except Exception as e:          # an auxiliary measurement must never cost the bench line (disk space, a missing tool)
            cli = {"error": f"{type(e).__name__}: {e}"[:300]}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u64 keys / u32 counts / f32 stats", "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nbytes + offs_host.nbytes),
                   "d2h_bytes_per_step": int(12 * nreads), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3},
           "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "table": {"capacity_slots": tinfo["capacity"], "distinct_kmers": tinfo["distinct"],
                     "load": round(tinfo["distinct"] / tinfo["capacity"], 3),
                     "stats_table_slots": qinfo["capacity"], "stats_table_kmers": qinfo["distinct"]},
           "device": {"sm_count": info["sm_count"], "hbm_total_gb": round(info["total_bytes"] / 1e9, 1)}}
    out["parity_checks"] = parity
    if cli is not None:
        out["cli_end_to_end"] = cli
    if phases is not None:
        out["multi_gpu"] = {"exchange": sc.exchange, "phases_ms_synced": phases, "partitions": sc.nparts, "partitions_per_rank": sc.lp,
                            "exchange_bins": sc.cbins, "replay_fold": sc.replay_fold}
        if routed is not None:
            out["multi_gpu"]["routed_statistics"] = routed
    if r2t is not None:
        out["reads_to_transcripts"] = r2t
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def run_reference_arm(args, config, world, W):
    """--impl reference: the reference's own CPU implementation of the step (count + coverage stats) on this box's
    host cores; each step is a bounded sample of the workload.  No GPU, none of our kernels."""
    read_len = args.read_len
    ns = min(args.cpu_sample_reads, 2 * args.pairs)
    # the sample is generated on the host with the test generator (same shape: reads from a weighted transcriptome)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synthdata
    rng = np.random.default_rng(SEED)
    txs = synthdata.transcriptome(rng, 2000)
    reads = synthdata.reads_from(rng, txs, ns, read_len, err=0.005, n_rate=0.001)
    reads = [r.ljust(read_len, b"N") for r in reads]
    recs = np.frombuffer(b"".join(r + b"\n" for r in reads), dtype=np.uint8)
    times = []
    kind, cores = "reference", 6
    for i in range(W + args.steps):
        v, dt, kind, cores = cpu_reference_run(recs, read_len, ns)
        if i >= W:
            times.append(dt)
    positions = 2 * ns * (read_len - K + 1)
    total = sum(times)
    value = positions * len(times) / total
    sample = (f"{ns} synthetic {read_len} bp reads per step (bounded sample of the workload): "
              f"fastaToKmerCoverageStats --kmers_from_reads + stats, --num_threads {cores}; host has {os.cpu_count()} cores")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": W, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64 keys / u32 counts / f32 stats", "data": "synthetic", "config": config,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
