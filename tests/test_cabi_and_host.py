"""CPU: the C-ABI library loads and exports every symbol include/trinity_gpu.h declares; host-side logic (entropy
table, reader restatements, formatting) without touching a GPU."""
import os
import re

import numpy as np
import pytest

import trinityrnaseq_b200 as tg
from trinityrnaseq_b200 import _lib
from oracle import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "trinity_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libtrinity_gpu.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.tg_version() >= 100


def test_no_cpu_fallback_without_gpu():
    if _lib.lib().tg_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(tg.TrinityGpuError) as e:
        tg.Context(0)
    assert e.value.code == _lib.TG_ERR_NOGPU and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_oracle():
    """the product path must never import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "trinityrnaseq_b200")
    for dirpath, _, files in os.walk(pkg):
        if any(part in ("build", "lib", "bin", "__pycache__") for part in dirpath.split(os.sep)):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "api.py" and "oracle" not in txt, (dirpath, f)


def test_entropy_table_matches_reference_expression():
    """tg_entropy_table (host side of R5) against the oracle's compute_entropy on an explicit window for EVERY count
    tuple, forward slot order and reverse-complement slot order, at the default and at other thresholds"""
    for thr in (1.5, 1.0, 1.9, 0.0):
        ok = np.zeros(26 ** 3, np.uint8)
        _lib.lib().tg_entropy_table(25, thr, ok.ctypes.data)
        n = 0
        for g in range(26):
            for a in range(26 - g):
                for t in range(26 - g - a):
                    c = 25 - g - a - t
                    e = orc.entropy("G" * g + "A" * a + "T" * t + "C" * c)
                    assert bool(ok[(g * 26 + a) * 26 + t]) == (not (np.float32(e) < np.float32(thr))), (g, a, t, c, thr)
                    n += 1
        assert n == 3276


def test_reader_restatements_edge_cases():
    txt = b"junk before\n>a b\tc\nAC GT\nac\tgt\n\n>b\n>c\nNNNN\n>d\nACGT"
    iw = orc.read_fasta_inchworm(txt)
    assert [(a, s) for _, a, s in iw] == [("a", "ACGTACGT"), ("b", ""), ("c", "NNNN"), ("d", "ACGT")]
    ds = orc.read_fasta_dnastream(txt)
    # the line after a header is always sequence (even '>c'); the unterminated last record is dropped
    assert ds == [(">a b\tc", "AC GTac\tgt"), (">b", ">cNNNN")]
    assert orc.read_fasta_dnastream(b">x\nACGT\nAC") == [(">x", "ACGT")]        # later unterminated line is lost
    assert orc.read_bundles(b">s_12 43 57\nacgtXacgt\n>s_13 1\nAC\nGT\n>s_14 2\nTTTT") == [
        (">s_12_43_57", "ACGTXACGT"), (">s_13_1", "ACGT"), (">s_14_2", "")]
    assert orc.format_read_name("> r name ") == ">_r_name_"


def test_stats_line_format():
    assert tg.format_stats_line("x/1", 67, np.float32(76.7308), np.float32(32.3844)) == "x/1\t67\t76.7308\t32.3844\tthread:0"
    nan = np.array([0xFFC00000], np.uint32).view(np.float32)[0]
    assert tg.format_stats_line("e", 1, np.float32(1), nan).split("\t")[3] == "-nan"
    assert tg.format_stats_line("s", 0, np.float32(0), np.float32(-0.0)).split("\t")[3] == "-0"
    assert tg.format_stats_line("b", 1234567, np.float32(1.33378e9), np.float32(2.30902e9)).split("\t")[2] == "1.33378e+09"


def test_packing_roundtrip():
    rng = np.random.default_rng(3)
    for _ in range(100):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, 25))
        assert tg.packed_to_kmer(tg.kmer_to_packed(s), 25) == s
    recs, offs = tg.records_from_sequences(["ACGT", "", "GG"])
    assert recs.tobytes() == b"ACGT\n\nGG\n" and offs.tolist() == [0, 5, 6, 9]


def test_sidecar_content_hash_cpp(tmp_path):
    """host/tg_sidecar.hpp (binary hand-off between `jellyfish dump` and the stats tool): the content hash is independent
    of the writer's chunking and notices single-byte edits and truncation -- compiled and run on the CPU."""
    import subprocess
    exe = tmp_path / "sidecar_hash_test"
    src = os.path.join(ROOT, "tests", "cpp", "sidecar_hash_test.cpp")
    inc = os.path.join(ROOT, "trinityrnaseq_b200", "host")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), src], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_host_fasta_readers_cpp_match_the_oracle_restatements(tmp_path):
    """The C++ readers the executables use (host/fasta_io.hpp) against the oracle's restatements of the reference's three
    readers, on the edge cases the survey probed (blank lines, blanks and lower case inside sequences, headers with
    tabs/spaces, '>' right after a header, unterminated last lines, CRLF-free multi-line records) and on a seeded random
    soup of such lines -- compiled and run on the CPU."""
    import subprocess
    exe = tmp_path / "readers_dump"
    src = os.path.join(ROOT, "tests", "cpp", "readers_dump.cpp")
    inc = os.path.join(ROOT, "trinityrnaseq_b200", "host")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), src], check=True)

    def dump(mode, data):
        f = tmp_path / "in.fa"
        f.write_bytes(data)
        r = subprocess.run([str(exe), mode, str(f)], capture_output=True, check=True)
        return [tuple(line.split(b"\x01")) for line in r.stdout.split(b"\n") if line]

    rng = np.random.default_rng(5)
    pieces = [b">r1 x\ty", b">r2", b">", b"> lead", b"ACGT", b"acgtn", b"AC GT", b"A\tC", b"", b" ", b">s_7 1 2", b"NNNN",
              b"ACGTXACGT", b"acgtxacgt tail", b"junk"]
    cases = [b"junk before\n>a b\tc\nAC GT\nac\tgt\n\n>b\n>c\nNNNN\n>d\nACGT",
             b">x\nACGT\nAC", b"", b"\n\n", b">only", b">only\n", b">a\nAC\n>b\nGT\n",
             b">s_12 43 57\nacgtXacgt\n>s_13 1\nAC\nGT\n>s_14 2\nTTTT"]
    for _ in range(60):
        n = int(rng.integers(1, 12))
        body = b"\n".join(pieces[int(i)] for i in rng.integers(0, len(pieces), n))
        cases.append(body + (b"\n" if rng.integers(0, 2) else b""))
    for data in cases:
        iw = [(a.encode(), s.encode()) for _, a, s in orc.read_fasta_inchworm(data)]
        assert dump("inchworm", data) == iw, data
        ds = orc.read_fasta_dnastream(data)
        got = dump("dnastream", data)
        assert [(g[0].split(b"\x02")[0], g[1]) for g in got] == [(n.encode(), s.encode()) for n, s in ds], data
        assert [g[0].split(b"\x02")[1] for g in got] == [orc.format_read_name(n).encode() for n, _ in ds], data
        bd = [(n.encode(), s.encode()) for n, s in orc.read_bundles(data)]
        assert dump("bundles", data) == bd, data


def test_jellyfish_sequence_file_parser_cpp(tmp_path):
    """host/seq_file.hpp: FASTA records are joined across line breaks (a k-mer may span them), FASTQ records give their
    second line, CR is dropped, text before the first header is ignored, batches are cut at record boundaries."""
    import subprocess
    exe = tmp_path / "seqfile_dump"
    src = os.path.join(ROOT, "tests", "cpp", "seqfile_dump.cpp")
    inc = os.path.join(ROOT, "trinityrnaseq_b200", "host")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), src], check=True)

    def parse(data, flush=1 << 30):
        f = tmp_path / "in.txt"
        f.write_bytes(data)
        r = subprocess.run([str(exe), str(f), str(flush)], capture_output=True, check=True)
        n, _, body = r.stdout.partition(b"\n")
        return int(n), body

    assert parse(b">a\nACGT\nAC\n>b desc\n\nGG\r\nTT\n>c\n>d\nNN")[1] == b"ACGTAC\nGGTT\n\nNN\n"
    assert parse(b"junk\nmore junk\n>a\nAC\n")[1] == b"AC\n"
    assert parse(b"")[1] == b"" and parse(b"\n\n")[1] == b"" and parse(b"no header at all\n")[1] == b""
    fq = b"@r1 x\nACGTN\n+\nIIIII\n@r2\nGG\r\n+r2\n@@\n@r3\nTTT"
    assert parse(fq)[1] == b"ACGTN\nGG\nTTT\n"            # a quality line may start with '@': records are 4 lines
    assert parse(b"\n\n" + fq)[1] == b"ACGTN\nGG\nTTT\n"
    # batches: many records, tiny threshold -> several flushes, same bytes, each cut after a terminator
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), int(rng.integers(0, 90)))) for _ in range(300)]
    fa = b"".join(b">s%d\n" % i + b"\n".join(s[j:j + 30] for j in range(0, len(s), 30)) + b"\n" for i, s in enumerate(seqs))
    want = b"".join(s + b"\n" for s in seqs)
    n1, body1 = parse(fa)
    n2, body2 = parse(fa, flush=500)
    assert body1 == want and body2 == want and n1 == 1 and n2 > 5
    fqs = b"".join(b"@q%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs))
    assert parse(fqs, flush=500)[1] == want


def test_parallel_fasta_parser_equals_serial(tmp_path):
    """host/par_fasta.hpp: chunked multi-threaded parsing == serial parsing, record for record, in file order, for FASTA
    texts with every reader quirk and chunk sizes down to one byte (tests/cpp/par_fasta_test.cpp)."""
    import subprocess
    exe = tmp_path / "par_fasta_test"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-pthread", "-I",
                    os.path.join(ROOT, "trinityrnaseq_b200", "host"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "cpp", "par_fasta_test.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "parallel parse == serial parse" in r.stdout


def test_float_formatting_equals_printf_g(tmp_path):
    """host/fmt_float.hpp (std::to_chars) prints what printf("%g") prints -- the reference's ostream default -- for 3 M floats
    of every kind, NaN signs and infinities included (tests/cpp/fmt_float_test.cpp)."""
    import subprocess
    exe = tmp_path / "fmt_float_test"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "trinityrnaseq_b200", "host"),
                    "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "fmt_float_test.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "0 differences" in r.stdout


def test_jellyfish_dump_blocks_and_threads(tmp_path):
    """`jellyfish dump` formats its records on a pool of threads, block by block (host/jellyfish_main.cpp); the text, the
    -L/-U filter and the .tgk sidecar (packed k-mers, counts, length and hash of the text) must be what one loop over the
    table writes.  No GPU involved: dump only reads the database file.  2.5 M k-mers = two blocks, every thread busy."""
    import struct
    import subprocess
    jf = os.path.join(ROOT, "trinityrnaseq_b200", "bin", "jellyfish")
    if not os.path.exists(jf):
        pytest.skip("executables not built")
    k, n = 25, 2_500_000
    rng = np.random.default_rng(77)
    keys = np.unique(rng.integers(0, 1 << 50, size=n, dtype=np.uint64))
    n = len(keys)
    cnts = rng.integers(10, 100, size=n).astype(np.uint32)          # two digits: fixed-width records, built with numpy below
    db = tmp_path / "t.jf"
    db.write_bytes(b"TGJF001\n" + struct.pack("<IIQ", k, 1, n) + b"\0" * (8 * 10002) + keys.tobytes() + cnts.tobytes())

    def expected(sel):
        kk, cc = keys[sel], cnts[sel]
        m = len(kk)
        lines = np.empty((m, 4 + k + 1), dtype=np.uint8)
        lines[:, 0] = ord(">")
        lines[:, 1] = ord("0") + cc // 10
        lines[:, 2] = ord("0") + cc % 10
        lines[:, 3] = ord("\n")
        shifts = (2 * (k - 1 - np.arange(k))).astype(np.uint64)
        codes = ((kk[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
        lines[:, 4:4 + k] = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
        lines[:, 4 + k] = ord("\n")
        return lines.tobytes()

    for lo, hi in ((1, None), (20, 80)):
        out = tmp_path / f"dump_{lo}.fa"
        cmd = [jf, "dump", "-L", str(lo)] + (["-U", str(hi)] if hi else []) + ["-o", str(out), str(db)]
        r = subprocess.run(cmd, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        sel = (cnts >= lo) & (cnts <= (hi if hi else 0xFFFFFFFF))
        want = expected(sel)
        assert out.read_bytes() == want
        side = (tmp_path / f"dump_{lo}.fa.tgk").read_bytes()
        magic, sk, _, sn, text_bytes, _ = struct.unpack("<8sIIQQQ", side[:40])
        assert magic == b"TGKMER1\n" and sk == k and sn == int(sel.sum()) and text_bytes == len(want)
        assert np.array_equal(np.frombuffer(side[40:40 + 8 * sn], dtype=np.uint64), keys[sel])
        assert np.array_equal(np.frombuffer(side[40 + 8 * sn:], dtype=np.uint32), cnts[sel])


def test_executables_fail_loudly_without_gpu(tmp_path):
    """no GPU -> the three drop-in executables leave with a non-zero status and say why (no CPU fallback anywhere)"""
    import subprocess
    if _lib.lib().tg_device_count() > 0:
        pytest.skip("a GPU is visible")
    bin_dir = os.path.join(ROOT, "trinityrnaseq_b200", "bin")
    if not os.path.exists(os.path.join(bin_dir, "jellyfish")):
        pytest.skip("executables not built")
    fa = os.path.join(ROOT, "tests", "golden", "reads.fa")
    bundles = os.path.join(ROOT, "tests", "golden", "bundles.fa")
    cmds = [[os.path.join(bin_dir, "fastaToKmerCoverageStats"), "--reads", fa, "--kmers_from_reads", fa],
            [os.path.join(bin_dir, "ReadsToTranscripts"), "-i", fa, "-f", bundles, "-o", str(tmp_path / "o")],
            [os.path.join(bin_dir, "jellyfish"), "count", "-m", "25", "-s", "1000", "-o", str(tmp_path / "m.jf"), fa]]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, timeout=120)
        assert r.returncode == 3, cmd
        assert b"no CPU fallback" in r.stderr and r.stdout == b""
