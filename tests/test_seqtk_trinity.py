"""seqtk-trinity drop-in (read prep in front of the k-mer path): stdout, exit code and error text against
the unmodified reference tool -- through committed goldens (tests/golden/seqtk, written by make_golden_seqtk.py from
oracle/_ref/seqtk-trinity) everywhere, and live against that binary, the reference's own fixtures
(trinity-plugins/seqtk-trinity/testing, exit codes of its test_seqtk_trinity.py) and random FASTQ where they exist.
CPU only: the tool is a text filter."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "seqtk")
OURS = os.path.join(ROOT, "trinityrnaseq_b200", "bin", "seqtk-trinity")
REF = os.path.join(ROOT, "oracle", "_ref", "seqtk-trinity")
REF_FIXTURES = "/root/reference/trinity-plugins/seqtk-trinity/testing"

pytestmark = pytest.mark.skipif(not os.path.exists(OURS), reason="executables not built")


def run(tool, args, path, via_stdin):
    if via_stdin:
        with open(path, "rb") as f:
            return subprocess.run([tool, "seq"] + args + ["-"], stdin=f, capture_output=True, timeout=120)
    return subprocess.run([tool, "seq"] + args + [path], capture_output=True, timeout=120)


def test_goldens():
    with open(os.path.join(GOLD, "expected.json")) as f:
        expect = json.load(f)
    assert len(expect) >= 80
    for e in expect:
        path = os.path.join(GOLD, e["input"])
        r = run(OURS, e["args"], path, e["stdin"])
        what = (e["input"], e["args"], e["stdin"])
        assert r.returncode == e["rc"], what
        with open(os.path.join(GOLD, e["stdout"]), "rb") as f:
            assert r.stdout == f.read(), what
        assert r.stderr.decode(errors="replace").replace(path, "<PATH>") == e["stderr"], what


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.isdir(REF_FIXTURES)), reason="reference tool / fixtures not here")
def test_reference_fixtures_live():
    """the reference's own test inputs, both read types: same bytes, same exit code (0 / 2 / 3 / 4)"""
    seen = set()
    for fq in sorted(os.listdir(REF_FIXTURES)):
        if not fq.endswith(".fq"):
            continue
        for rt in ("1", "2"):
            for extra in ([], ["-r"]):
                args = ["-A", "-R", rt] + extra
                a = run(OURS, args, os.path.join(REF_FIXTURES, fq), False)
                b = run(REF, args, os.path.join(REF_FIXTURES, fq), False)
                assert (a.returncode, a.stdout, a.stderr) == (b.returncode, b.stdout, b.stderr), (fq, args)
                seen.add(b.returncode)
    assert {0, 2, 3, 4} <= seen


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tool not here")
def test_random_fastq_live(tmp_path):
    """a few thousand random records (names of every format, CR LF, wrapped lines, blank lines, IUPAC letters, gzip) and
    a 30 MB file that crosses the tool's 4 MiB input window many times"""
    import gzip
    rng = np.random.default_rng(99)
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMSWBDHVU-", dtype=np.uint8)

    def records(n, lo, hi):
        out = []
        for i in range(n):
            L = int(rng.integers(lo, hi))
            s = alphabet[rng.integers(0, len(alphabet), L)].tobytes()
            q = bytes(rng.integers(33, 75, L).astype(np.uint8))
            name = [b"r%d/1", b"r%d 1:N:0:7", b"r%d", b"r%d_forward/1", b"r%d\tx y", b"r%d/1 extra"][i % 6] % i
            eol = b"\r\n" if i % 11 == 0 else b"\n"
            if i % 7 == 0 and L > 20:
                s = s[:L // 2] + eol + s[L // 2:]
            out.append(b"@" + name + eol + s + eol + b"+" + (name if i % 5 == 0 else b"") + eol + q + eol + (b"\n" if i % 13 == 0 else b""))
        return b"".join(out)

    small = tmp_path / "small.fq"
    small.write_bytes(records(3000, 1, 300))
    big = tmp_path / "big.fq"
    big.write_bytes(records(100000, 100, 200))
    gz = tmp_path / "small.fq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(small.read_bytes())
    for path in (small, big, gz):
        for args in (["-A", "-R", "1"], ["-A", "-R", "1", "-r"], ["-R", "1", "-l", "60"], ["-A", "-R", "1", "-q", "15", "-U"]):
            for via_stdin in (False, True):
                a = run(OURS, args, str(path), via_stdin)
                b = run(REF, args, str(path), via_stdin)
                assert a.returncode == b.returncode == 0, (path.name, args, b.stderr[:200])
                assert a.stdout == b.stdout, (path.name, args, via_stdin)
