#!/usr/bin/env python
"""Golden md5 vectors for BASELINE-shaped slices (VERDICT r01 'next' 1c) and the reference's own fixtures (1b).

Run in the build container (needs /root/reference and oracle/_ref built by oracle/Makefile.ref):
    python tests/golden/make_golden_big.py
It (a) copies the reference's shipped INPUT fixtures (data, not code) into tests/golden/ref/, (b) regenerates the
seeded synthetic slices with tests/synthdata.py, (c) runs the UNMODIFIED reference binaries on all of them and stores
the md5 of the normalised outputs in tests/golden/big_slices.json.  The GPU tests regenerate the same inputs from
the same seeds, run the GPU executables and compare md5s -- no reference binary and no /root/reference at run time.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bigslices  # noqa: E402

REF = "/root/reference"
REFBIN = os.path.join(ROOT, "oracle", "_ref")
ENV = dict(os.environ, LC_ALL="C", OMP_NUM_THREADS="1")


def main():
    dst = os.path.join(HERE, "ref")
    os.makedirs(dst, exist_ok=True)
    for src in ("trinity_ext_sample_data/__regression_tests/test_GraphFromFasta/both.fa.gz",
                "trinity_ext_sample_data/__regression_tests/test_GraphFromFasta/inchworm.K25.L25.fa.gz",
                "sample_data/test_Trinity_Assembly/reads.left.fa.gz",
                "sample_data/test_Trinity_Assembly/reads.right.fa.gz"):
        shutil.copyfile(os.path.join(REF, src), os.path.join(dst, os.path.basename(src)))
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, spec in bigslices.SLICES.items():
            files = bigslices.materialise(name, td)
            res = {}
            if spec["tool"] == "stats":
                for mode in spec["modes"]:
                    cmd = [os.path.join(REFBIN, "fastaToKmerCoverageStats"), "--reads", files["reads"], "--kmers_from_reads",
                           files["reads"], "--kmer_size", "25", "--num_threads", "1", "--" + mode]
                    r = subprocess.run(cmd, capture_output=True, env=ENV, check=True)
                    res[mode] = bigslices.stats_md5(r.stdout, td, spec.get("keep_header", False))
            else:
                for mode in spec["modes"]:
                    o = os.path.join(td, "ref.out")
                    cmd = [os.path.join(REFBIN, "ReadsToTranscripts"), "-i", files["reads"], "-f", files["bundles"], "-o", o,
                           "-t", "1", "-max_mem_reads", "50000000", "-p", "10"] + (["-strand"] if mode == "strand" else [])
                    subprocess.run(cmd, capture_output=True, env=ENV, check=True)
                    res[mode] = {"md5": bigslices.r2t_md5(o, td), "rcts": open(o + ".rcts.out").read().strip()}
            res["inputs_md5"] = {k: hashlib.md5(open(v, "rb").read()).hexdigest() for k, v in files.items()}
            out[name] = res
            print(name, res, flush=True)
    with open(os.path.join(HERE, "big_slices.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
