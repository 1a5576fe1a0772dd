"""GPU: the drop-in executables on BASELINE-shaped slices and on the reference's own fixtures, compared with the md5 of the
UNMODIFIED reference binaries' normalised output (tests/golden/big_slices.json, made by tests/golden/make_golden_big.py in
the build container; SURVEY App. B for the reference fixtures).  Inputs are regenerated here from the same seeds (their
md5 is checked too), so nothing from /root/reference is needed at run time.

  c2_2x100_1M      1 M reads of configs[1]'s shape through count (log path) + statistics, --DS and --SS
  c3_2x150_700k    configs[2]'s read shape with 1 MB partitions: > 512 partitions => coarse log + refine path
  c4_r2t_300k      configs[3]'s shape: readsToComponents.out.sort vs `ReadsToTranscripts -t 1`, DS and -strand
  ref_*            __regression_tests/test_GraphFromFasta (9,689 k-mers shared between bundles: rule R4 at scale) and
                   the 61,150-read sample (fastaToKmerCoverageStats.cpp:122-172, ReadsToTranscripts.cc:146-169,216-297)
"""
import hashlib
import json
import os
import subprocess

import pytest

import bigslices

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "trinityrnaseq_b200", "bin")
with open(os.path.join(bigslices.GOLD, "big_slices.json")) as f:
    EXPECT = json.load(f)

CASES = [(n, m) for n, s in bigslices.SLICES.items() for m in s["modes"]]


@pytest.fixture(scope="module")
def slice_files(tmp_path_factory):
    cache = {}

    def get(name):
        if name not in cache:
            td = str(tmp_path_factory.mktemp(name))
            files = bigslices.materialise(name, td)
            for k, p in files.items():      # same bytes the reference binaries saw
                assert hashlib.md5(open(p, "rb").read()).hexdigest() == EXPECT[name]["inputs_md5"][k], (name, k)
            cache[name] = (files, td)
        return cache[name]
    return get


@pytest.mark.parametrize("name,mode", CASES)
def test_slice_matches_reference_md5(slice_files, name, mode):
    spec = bigslices.SLICES[name]
    files, td = slice_files(name)
    env = dict(bigslices.ENV, **spec["env"])
    if spec["tool"] == "stats":
        r = subprocess.run([os.path.join(BIN, "fastaToKmerCoverageStats"), "--reads", files["reads"], "--kmers_from_reads",
                            files["reads"], "--kmer_size", "25", "--num_threads", "6", "--" + mode], capture_output=True, env=env,
                           timeout=900)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        got = bigslices.stats_md5(r.stdout, td, spec.get("keep_header", False))
        assert got == EXPECT[name][mode]
    else:
        out = os.path.join(td, f"ours_{mode}.out")
        cmd = [os.path.join(BIN, "ReadsToTranscripts"), "-i", files["reads"], "-f", files["bundles"], "-o", out, "-t", "8",
               "-max_mem_reads", "50000000", "-p", "10"] + (["-strand"] if mode == "strand" else [])
        r = subprocess.run(cmd, capture_output=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        assert open(out + ".rcts.out").read().strip() == EXPECT[name][mode]["rcts"]
        assert bigslices.r2t_md5(out, td) == EXPECT[name][mode]["md5"]
    if (name, mode) in bigslices.SURVEY_MD5:     # the survey's own vectors, verbatim
        exp = EXPECT[name][mode]
        assert (exp if isinstance(exp, str) else exp["md5"]) == bigslices.SURVEY_MD5[(name, mode)]


def test_jellyfish_dump_feeds_stats_like_kmers_from_reads(slice_files, tmp_path):
    """SURVEY §8c cross-check at size: our `jellyfish count | dump -L 1` fed to our stats tool (--kmers) must give the
    reference's --kmers_from_reads md5 (no read of the slice has length exactly K, so S3's skip cannot fire)."""
    name = "c2_2x100_1M"
    files, td = slice_files(name)
    jf = str(tmp_path / "mer_counts.jf")
    dump = str(tmp_path / "kmers.fa")
    jelly = os.path.join(BIN, "jellyfish")
    r = subprocess.run([jelly, "count", "-t", "8", "-m", "25", "-s", "100000000", "-o", jf, "--canonical", files["reads"]],
                       capture_output=True, env=bigslices.ENV, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    with open(dump, "wb") as f:
        r = subprocess.run([jelly, "dump", "-L", "1", jf], stdout=f, stderr=subprocess.PIPE, env=bigslices.ENV, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    for sidecar in ("0", "1"):        # through the binary hand-off and through the text parser
        env = dict(bigslices.ENV, TRINITY_GPU_NO_SIDECAR=sidecar)
        r = subprocess.run([os.path.join(BIN, "fastaToKmerCoverageStats"), "--reads", files["reads"], "--kmers", dump,
                            "--kmer_size", "25", "--num_threads", "6", "--DS"], capture_output=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        assert bigslices.stats_md5(r.stdout, td) == EXPECT[name]["DS"]
