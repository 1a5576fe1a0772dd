"""CPU: the C++ drop-ins of the two Perl helpers that turn coverage statistics into the normalised read set
(util/support_scripts/nbkc_merge_left_right_stats.pl, nbkc_normalize.pl) against the committed
outputs of the UNMODIFIED scripts (tests/golden/make_golden_nbkc.py), byte for byte -- including Perl's rand() stream
after srand(12345), its numification of "-nan" / "-0" / integers, "%.1f" / "NaN" formatting and split() semantics.
When perl and the reference tree are present (the build container) the scripts are also run live on fresh random tables."""
import os
import random
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "trinityrnaseq_b200", "bin")
MERGE = os.path.join(BIN, "nbkc_merge_left_right_stats.pl")
NORM = os.path.join(BIN, "nbkc_normalize.pl")
REF = "/root/reference/util/support_scripts"
CASES = {"a": ["--max_cov", "50", "--min_cov", "1", "--max_CV", "10000"],
         "b": ["--max_cov", "5", "--min_cov", "2", "--max_CV", "1"],
         "c": ["--max_cov=200", "--max_CV=10000"]}


def gold(name):
    with open(os.path.join(GOLD, name), "rb") as f:
        return f.read()


def run(cmd):
    return subprocess.run(cmd, capture_output=True, timeout=120)


def test_merge_left_right_matches_reference_script():
    r = run([MERGE, "--left", os.path.join(GOLD, "nbkc_left.stats"), "--right", os.path.join(GOLD, "nbkc_right.stats"), "--sorted"])
    assert r.returncode == 0 and r.stdout == gold("nbkc_pairs.expected")
    assert b"-done opening files." in r.stderr
    # unpaired entries without --sorted are fatal, like the script's die
    r = run([MERGE, "--left", os.path.join(GOLD, "nbkc_left.stats"), "--right", os.path.join(GOLD, "nbkc_right.stats")])
    assert r.returncode != 0 and b"core accs are not equivalent" in r.stderr
    assert run([MERGE]).returncode == 255


@pytest.mark.parametrize("tag", sorted(CASES))
def test_normalize_matches_reference_script(tag):
    r = run([NORM, "--stats_file", os.path.join(GOLD, "nbkc_pairs.expected")] + CASES[tag])
    assert r.returncode == 0
    assert r.stdout == gold(f"nbkc_selected_{tag}.expected")          # the same rand() draws in the same order
    assert r.stderr == gold(f"nbkc_selected_{tag}.stderr")


def test_normalize_single_end_table_and_cli_errors(tmp_path):
    r = run([NORM, "--stats_file", os.path.join(GOLD, "nbkc_left.stats"), "--max_cov", "30", "--min_cov", "1", "--max_CV", "100"])
    assert r.returncode == 0 and r.stdout == gold("nbkc_selected_single.expected")
    assert run([NORM]).returncode == 255 and b"--stats_file" in run([NORM]).stderr
    assert run([NORM, "--stats_file", str(tmp_path / "missing"), "--max_cov", "5", "--max_CV", "1"]).returncode != 0
    empty = tmp_path / "empty.stats"
    empty.write_text("acc\tmedian_cov\tmean_cov\tstdev\ttid\n")
    r = run([NORM, "--stats_file", str(empty), "--max_cov", "5", "--max_CV", "1"])
    assert r.returncode != 0 and b"no reads made it" in r.stderr
    bad = tmp_path / "bad.stats"
    bad.write_text("acc\tmedian_cov\tmean_cov\tstdev\ttid\nr/1\t3\t3.5\n")
    assert run([NORM, "--stats_file", str(bad), "--max_cov", "5", "--max_CV", "1"]).returncode != 0


@pytest.mark.skipif(not (shutil.which("perl") and os.path.isdir(REF)), reason="needs perl and the reference tree")
def test_live_against_the_perl_scripts(tmp_path):
    rnd = random.Random(20251017)
    for trial in range(4):
        files = []
        for side in (1, 2):
            rows = []
            for i in range(rnd.randint(1, 900)):
                if rnd.random() < 0.1:
                    continue
                med = rnd.choice([0, 1, 2, 3, 7, 20, 199, 200, 201, 65535, 4294967295])
                mean = "%g" % (med * rnd.uniform(0.5, 2.0))
                sd = rnd.choice(["-nan", "nan", "-0", "0", "-0.0", "inf", "%g" % rnd.uniform(0, 1e7), "%g" % rnd.uniform(0, 3)])
                name = rnd.choice(["r%04d/%d", "r%04d_%d", "r%04d/%d"]) % (i, side)
                rows.append((name, str(med), mean, sd, "thread:%d" % rnd.randint(0, 5)))
            rows.sort(key=lambda r: r[0].encode())
            p = tmp_path / f"t{trial}_{side}.stats"
            p.write_text("acc\tmedian_cov\tmean_cov\tstdev\ttid\n" + "".join("\t".join(r) + "\n" for r in rows))
            files.append(str(p))
        a = run(["perl", REF + "/nbkc_merge_left_right_stats.pl", "--left", files[0], "--right", files[1], "--sorted"])
        b = run([MERGE, "--left", files[0], "--right", files[1], "--sorted"])
        assert a.returncode == b.returncode == 0 and a.stdout == b.stdout
        pairs = tmp_path / f"p{trial}.stats"
        pairs.write_bytes(a.stdout)
        for args in (["--max_cov", "30", "--min_cov", "1", "--max_CV", "2"], ["--max_cov", "200", "--min_cov", "3", "--max_CV", "10000"]):
            for table in (str(pairs), files[0]):
                x = run(["perl", REF + "/nbkc_normalize.pl", "--stats_file", table] + args)
                y = run([NORM, "--stats_file", table] + args)
                assert (x.returncode == 0) == (y.returncode == 0)
                assert x.stdout == y.stdout
                if x.returncode == 0:
                    assert x.stderr == y.stderr


def test_sign_of_a_zero_sum_follows_perl(tmp_path):
    """"%.1f" of ($left + $right) / 2 for every pair of zero spellings, as printed by perl 5.38 (measured once, see
    host/perl_compat.hpp perl_add): the sum is -0.0 only when no integer coercion intervened."""
    z = ["-0", "-0.0", "0", "0.0", "-0e0", "+0", "-0.00"]
    table = ["0.0 -0.0 0.0 0.0 0.0 0.0 -0.0",      # left -0
             "0.0 -0.0 0.0 0.0 -0.0 0.0 -0.0",     # left -0.0
             "0.0 0.0 0.0 0.0 0.0 0.0 0.0",        # left 0
             "0.0 0.0 0.0 0.0 0.0 0.0 0.0",        # left 0.0
             "0.0 -0.0 0.0 0.0 0.0 0.0 -0.0",      # left -0e0
             "0.0 0.0 0.0 0.0 0.0 0.0 0.0",        # left +0
             "0.0 -0.0 0.0 0.0 -0.0 0.0 -0.0"]     # left -0.00
    left = tmp_path / "l.stats"
    right = tmp_path / "r.stats"
    hdr = "acc\tmedian_cov\tmean_cov\tstdev\ttid\n"
    left.write_text(hdr + "".join("p%02d/1\t1\t1\t%s\tt\n" % (i * 7 + j, a) for i, a in enumerate(z) for j, _ in enumerate(z)))
    right.write_text(hdr + "".join("p%02d/2\t1\t1\t%s\tt\n" % (i * 7 + j, b) for i, _ in enumerate(z) for j, b in enumerate(z)))
    r = run([MERGE, "--left", str(left), "--right", str(right)])
    assert r.returncode == 0
    got = [l.split(b"\t")[-1].decode() for l in r.stdout.split(b"\n")[1:] if l]
    assert got == [v for row in table for v in row.split()]
