#!/usr/bin/env python
"""Whole-process run of the drop-in statistics executable on a configs[1]-shaped FASTA file with TRINITY_GPU_TRACE=1: prints
the tool's own phase clock (stderr) and the wall time.   python tools/exp_cli.py [--reads N] [--gpus 0,1]"""
import argparse, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import trinityrnaseq_b200 as tg
from bench import make_transcriptome, make_bundles, write_sample_fasta, SEED, K
ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=20_000_000)
ap.add_argument("--read-len", type=int, default=100)
ap.add_argument("--gpus", default="")
ap.add_argument("--jellyfish", action="store_true", help="also time jellyfish count / dump -L 1 / dump -L 2 on the file")
ap.add_argument("--r2t", action="store_true", help="also time the ReadsToTranscripts executable on the file (bundles cut from the transcriptome)")
ap.add_argument("--no-stats", action="store_true")
ap.add_argument("--hold-gpu-gb", type=float, default=0, help="device memory the parent keeps allocated while the tool runs")
ap.add_argument("--hold-pinned-gb", type=float, default=0, help="pinned host memory the parent keeps while the tool runs")
ap.add_argument("--sleep", type=float, default=0, help="seconds between runs (the driver finishes a process's teardown in the background)")
ap.add_argument("--variants", default="", help="environment variants, e.g. 'A=1,B=2;;C=3' (an empty one = defaults)")
a = ap.parse_args()
ctx = tg.Context(0)
tx, tx_offs, tx_cum = make_transcriptome(20000, SEED)
d_recs, nbytes = ctx.synth_reads_dev(tx, tx_offs, tx_cum, a.reads // 2, a.read_len, seed=SEED)
recs = ctx.d2h(d_recs, nbytes, np.uint8).copy()
ctx.dev_free(d_recs)        # (the context stays open, as the bench's does while it times the executable)
held = []
if a.hold_gpu_gb:
    held.append(ctx.dev_alloc(int(a.hold_gpu_gb * 2**30)))
if a.hold_pinned_gb:
    held.append(ctx.pinned((int(a.hold_pinned_gb * 2**30),), np.uint8))
exe = os.path.join(ROOT, "trinityrnaseq_b200", "bin", "fastaToKmerCoverageStats")
with tempfile.TemporaryDirectory() as td:
    fa = os.path.join(td, "reads.fa")
    write_sample_fasta(fa, recs, a.read_len, a.reads)
    env = dict(os.environ, TRINITY_GPU_TRACE="1")
    if a.gpus:
        env["TRINITY_GPUS"] = a.gpus
    variants = [{}, {}, {}]
    if a.variants:
        variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in a.variants.split(";")]
    if a.no_stats:
        variants = []
    for rep, extra in enumerate(variants):
        time.sleep(a.sleep)
        run_env = dict(env, **extra)
        print("env", extra)
        print(f"launch at wall {time.time():.3f}")
        t0 = time.perf_counter()
        with open(os.path.join(td, "out.stats"), "wb") as so:
            r = subprocess.run([exe, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", str(K), "--DS"], stdout=so,
                               stderr=subprocess.PIPE, env=run_env)
        dt = time.perf_counter() - t0
        print(f"reaped at wall {time.time():.3f}")
        print(f"rep {rep}: rc {r.returncode}, wall {dt:.3f} s, fasta {os.path.getsize(fa) / 1e9:.2f} GB, out "
              f"{os.path.getsize(os.path.join(td, 'out.stats')) / 1e9:.2f} GB")
        print("\n".join(l for l in r.stderr.decode().splitlines() if "[trace]" in l))

    if a.jellyfish:
        jf = os.path.join(ROOT, "trinityrnaseq_b200", "bin", "jellyfish")
        db = os.path.join(td, "mer.jf")
        for rep in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([jf, "count", "-t", "16", "-m", str(K), "-s", "1000000000", "--canonical", "-o", db, fa], stderr=subprocess.PIPE)
            t1 = time.perf_counter()
            print(f"jellyfish count rep {rep}: rc {r.returncode}, {t1 - t0:.3f} s, db {os.path.getsize(db) / 1e9:.2f} GB")
        for L in (1, 2):
            t0 = time.perf_counter()
            r = subprocess.run([jf, "dump", "-L", str(L), "-o", os.path.join(td, "dump.fa"), db], stderr=subprocess.PIPE)
            t1 = time.perf_counter()
            print(f"jellyfish dump -L {L}: rc {r.returncode}, {t1 - t0:.3f} s, text {os.path.getsize(os.path.join(td, 'dump.fa')) / 1e9:.2f} GB")

    if a.r2t:
        brecs, boffs, ncontigs = make_bundles(tx, tx_offs, SEED + 1)
        bf = os.path.join(td, "bundles.fa")
        with open(bf, "wb") as f:
            for i in range(len(boffs) - 1):
                f.write(b">s_%d 10\n" % i)
                f.write(brecs[int(boffs[i]):int(boffs[i + 1])].tobytes())
        r2t = os.path.join(ROOT, "trinityrnaseq_b200", "bin", "ReadsToTranscripts")
        for rep in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([r2t, "-i", fa, "-f", bf, "-o", os.path.join(td, "r2t.out"), "-t", "16", "-max_mem_reads", "50000000", "-p", "10"],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t0
            print(f"ReadsToTranscripts rep {rep}: rc {r.returncode}, {dt:.3f} s = {a.reads / dt / 1e6:.2f} M reads/s, out "
                  f"{os.path.getsize(os.path.join(td, 'r2t.out')) / 1e9:.2f} GB, {ncontigs} contigs")
