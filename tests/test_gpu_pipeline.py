"""GPU: the reference's UNMODIFIED Perl driver of the normalisation stage (util/insilico_read_normalization.pl, staged
verbatim under oracle/_ref/trinity_home by oracle/Makefile.ref) run end to end on the reference's own sample reads with
the drop-in executables -- `jellyfish` first on $PATH, `fastaToKmerCoverageStats` at $TRINITY_HOME/Inchworm/bin -- exactly
as Trinity calls it: --pairs_together --PARALLEL_STATS --max_cov 200 (Trinity:3444-3449, :214).  --PARALLEL_STATS makes
the driver run the left and the right statistics as two concurrent processes sharing the one GPU
(util/insilico_read_normalization.pl:853-855, 990-1015).

Checked: (1) the selected accession list and the normalised read files are byte-identical to the same run with the
UNMODIFIED reference statistics tool (oracle/_ref/fastaToKmerCoverageStats) in that place; (2) the same again with the
C++ drop-ins for nbkc_merge_left_right_stats.pl / nbkc_normalize.pl instead of the Perl scripts.
(insilico_read_normalization.pl:588-654 run_jellyfish, :816-898 generate_stats_files, :951-985 run_nkbc_pairs_together.)
"""
import gzip
import hashlib
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "trinityrnaseq_b200", "bin")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
HOME0 = os.path.join(REFDIR, "trinity_home")
GOLD = os.path.join(ROOT, "tests", "golden", "ref")


def _home(tmp, name, stats_tool, cxx_nbkc):
    """a TRINITY_HOME whose statistics tool (and optionally nbkc helpers) are the ones under test"""
    home = os.path.join(tmp, name)
    shutil.copytree(HOME0, home)
    os.symlink(stats_tool, os.path.join(home, "Inchworm", "bin", "fastaToKmerCoverageStats"))
    if cxx_nbkc:
        for s in ("nbkc_merge_left_right_stats.pl", "nbkc_normalize.pl"):
            dst = os.path.join(home, "util", "support_scripts", s)
            os.unlink(dst)
            shutil.copy(os.path.join(BIN, s), dst)
    return home


def _run(home, outdir, left, right):
    env = dict(os.environ, LC_ALL="C", PATH=BIN + os.pathsep + os.environ["PATH"])      # our jellyfish first on $PATH
    cmd = ["perl", os.path.join(home, "util", "insilico_read_normalization.pl"), "--seqType", "fa", "--JM", "1G", "--max_cov", "200",
           "--min_cov", "1", "--CPU", "4", "--output", outdir, "--max_CV", "10000", "--left", left, "--right", right,
           "--pairs_together", "--PARALLEL_STATS", "--no_cleanup"]        # keep tmp_normalized_reads/ (the .accs list)
    r = subprocess.run(cmd, capture_output=True, env=env, timeout=1200)
    assert r.returncode == 0, (r.stdout.decode()[-2000:], r.stderr.decode()[-4000:])
    out = {}
    for root, _, files in os.walk(outdir):
        for f in files:
            if f.endswith(".accs") or (".normalized_" in f and not f.endswith(".ok")):
                p = os.path.join(root, f)
                if os.path.islink(p):
                    p = os.path.realpath(p)
                out[f] = hashlib.md5(open(p, "rb").read()).hexdigest()
    return out


@pytest.mark.skipif(not os.path.exists(os.path.join(HOME0, "util", "insilico_read_normalization.pl")),
                    reason="oracle/_ref/trinity_home not staged (oracle/Makefile.ref needs /root/reference once)")
def test_unmodified_normalisation_driver_with_dropins(tmp_path):
    tmp = str(tmp_path)
    left, right = os.path.join(tmp, "reads.left.fa"), os.path.join(tmp, "reads.right.fa")
    for src, dst in (("reads.left.fa.gz", left), ("reads.right.fa.gz", right)):
        with open(dst, "wb") as f:
            f.write(gzip.open(os.path.join(GOLD, src)).read())
    ours = _run(_home(tmp, "home_ours", os.path.join(BIN, "fastaToKmerCoverageStats"), False), os.path.join(tmp, "out_ours"), left, right)
    ref = _run(_home(tmp, "home_ref", os.path.join(REFDIR, "fastaToKmerCoverageStats"), False), os.path.join(tmp, "out_ref"), left, right)
    cxx = _run(_home(tmp, "home_cxx", os.path.join(BIN, "fastaToKmerCoverageStats"), True), os.path.join(tmp, "out_cxx"), left, right)
    accs = [k for k in ref if k.endswith(".accs")]
    norm = [k for k in ref if ".normalized_" in k]
    assert accs and len(norm) == 2, sorted(ref)
    for k in accs + norm:
        assert ours.get(k) == ref[k], ("GPU statistics tool changed " + k, ours, ref)
        assert cxx.get(k) == ref[k], ("C++ nbkc helpers changed " + k, cxx, ref)
    # the sorted statistics files differ only in the thread column of the reference tool: columns 1-4 are compared by the
    # big-slice tests; here the selections are the contract
