"""GPU: the three drop-in executables against the committed outputs of the unmodified reference binaries
(tests/golden/*.expected, produced with -t 1 / --num_threads 1 so that line order is deterministic) -- byte for byte."""
import os
import subprocess

import numpy as np
import pytest

import trinityrnaseq_b200 as tg
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "trinityrnaseq_b200", "bin")
ENV = dict(os.environ, LC_ALL="C")


def gold(name):
    with open(os.path.join(GOLD, name), "rb") as f:
        return f.read()


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, env=ENV, timeout=300, **kw)


@pytest.mark.parametrize("tag,mode", [("", "DS"), ("", "SS"), ("_nonl", "DS"), ("_nonl", "SS")])
def test_stats_kmers_from_reads(tag, mode):
    fa = os.path.join(GOLD, f"reads{tag}.fa")
    r = run([os.path.join(BIN, "fastaToKmerCoverageStats"), "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "25",
             "--num_threads", "6", "--" + mode])
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == gold(f"stats{tag}_{mode}.expected")
    assert b"STATS_GENERATION_TIME" in r.stderr


def test_stats_variants():
    fa = os.path.join(GOLD, "reads.fa")
    exe = os.path.join(BIN, "fastaToKmerCoverageStats")
    r = run([exe, "--reads", fa, "--kmers_from_reads", fa, "--capture_coverage_info"])
    assert r.returncode == 0 and r.stdout == gold("stats_capture.expected")
    r = run([exe, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "21"])
    assert r.returncode == 0 and r.stdout == gold("stats_k21.expected")
    r = run([exe, "--reads", fa, "--kmers", os.path.join(GOLD, "kmers_L2.fa"), "--kmer_size", "25", "--DS"])
    assert r.returncode == 0 and r.stdout == gold("stats_kmers_L2.expected")
    assert b"is not of length: 25" in r.stderr
    # CLI errors (fastaToKmerCoverageStats.cpp:63-82): usage -> 1, k < 20 -> 2, unreadable file -> 1
    assert run([exe, "--reads", fa]).returncode == 1
    assert run([exe, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "19"]).returncode == 2
    assert run([exe, "--reads", "/nonexistent.fa", "--kmers_from_reads", fa]).returncode == 1


@pytest.mark.parametrize("tag,mode", [("", "ds"), ("", "strand"), ("_nonl", "ds"), ("_nonl", "strand")])
def test_reads_to_transcripts(tmp_path, tag, mode):
    out = tmp_path / "r2c.out"
    cmd = [os.path.join(BIN, "ReadsToTranscripts"), "-i", os.path.join(GOLD, f"reads{tag}.fa"), "-f",
           os.path.join(GOLD, "bundles.fa"), "-o", str(out), "-t", "8", "-max_mem_reads", "50000000", "-p", "0"]
    if mode == "strand":
        cmd.append("-strand")
    r = run(cmd)
    assert r.returncode == 0, r.stderr.decode()
    assert out.read_bytes() == gold(f"r2t{tag}_{mode}.expected")
    assert (tmp_path / "r2c.out.rcts.out").read_bytes() == gold(f"r2t{tag}_{mode}.expected.rcts.out")


def test_reads_to_transcripts_chunks_and_cli(tmp_path):
    out = tmp_path / "o"
    exe = os.path.join(BIN, "ReadsToTranscripts")
    r = run([exe, "-i", os.path.join(GOLD, "reads.fa"), "-f", os.path.join(GOLD, "bundles.fa"), "-o", str(out),
             "-max_mem_reads", "100", "-p", "10"])
    assert r.returncode == 0
    assert out.read_bytes() == gold("r2t_p10_chunk100.expected")
    assert (tmp_path / "o.rcts.out").read_bytes() == gold("r2t_p10_chunk100.expected.rcts.out")
    assert run([exe, "-bogus", "1"]).returncode == 255          # exit(-1) on an unknown argument
    assert run([exe, "-h"]).returncode == 255


def test_jellyfish_count_dump_histo(tmp_path):
    jf = os.path.join(BIN, "jellyfish")
    fa = os.path.join(GOLD, "reads.fa")
    v = run([jf, "--version"])
    assert v.returncode == 0 and v.stdout.split()[1].startswith(b"2.")
    # what jellyfish sees: FASTA records, line breaks inside a record do not break k-mers, blanks do
    seqs = []
    for line in gold("reads.fa").split(b"\n"):
        if line.startswith(b">"):
            seqs.append(b"")
        elif seqs:
            seqs[-1] += line
    recs, _ = tg.records_from_sequences(seqs)
    for canonical in (True, False):
        db = tmp_path / f"mer_{int(canonical)}.jf"
        cmd = [jf, "count", "-t", "4", "-m", "25", "-s", "100000000", "-o", str(db)] + (["--canonical"] if canonical else []) + [fa]
        r = run(cmd)
        assert r.returncode == 0, r.stderr.decode()
        for L in (1, 2):
            keys, cnts = orc.jf_count(recs, 25, canonical, L)
            expect = "".join(">%d\n%s\n" % (c, tg.packed_to_kmer(k, 25)) for k, c in zip(keys, cnts)).encode()
            d = run([jf, "dump", "-L", str(L), str(db)])
            assert d.returncode == 0 and d.stdout == expect
        keys, cnts = orc.jf_count(recs, 25, canonical, 1)
        bins = orc.jf_histo(cnts)
        expect = "".join("%d %d\n" % (c, bins[c]) for c in range(1, 10002) if bins[c]).encode()
        h = run([jf, "histo", "-t", "4", "-o", str(tmp_path / "h.txt"), str(db)])
        assert h.returncode == 0 and (tmp_path / "h.txt").read_bytes() == expect
        c = run([jf, "dump", "-c", "-L", "3", str(db)])
        k3, c3 = orc.jf_count(recs, 25, canonical, 3)
        assert c.stdout == "".join("%s %d\n" % (tg.packed_to_kmer(k, 25), n) for k, n in zip(k3, c3)).encode()
    assert run([jf, "count", "-m", "25", fa]).returncode != 0          # -s is required
    assert run([jf, "dump", str(tmp_path / "missing.jf")]).returncode != 0


def test_pipeline_dump_into_stats(tmp_path):
    """The normalisation pipeline's hand-off (util/insilico_read_normalization.pl:617-846): jellyfish count --canonical
    | dump -L 1 -> fastaToKmerCoverageStats --kmers must equal counting the reads directly, when no read has exactly
    k bases (SURVEY §8c cross-check)."""
    jf = os.path.join(BIN, "jellyfish")
    stats = os.path.join(BIN, "fastaToKmerCoverageStats")
    entries = orc.read_fasta_inchworm(gold("reads.fa"))
    fa = tmp_path / "r.fa"
    fa.write_text("".join(">%s\n%s\n" % (h, s) for h, _, s in entries if len(s) != 25))
    assert run([jf, "count", "-t", "2", "-m", "25", "-s", "1000000", "--canonical", "-o", str(tmp_path / "m.jf"), str(fa)]).returncode == 0
    d = run([jf, "dump", "-L", "1", str(tmp_path / "m.jf")])
    (tmp_path / "k.fa").write_bytes(d.stdout)
    a = run([stats, "--reads", str(fa), "--kmers", str(tmp_path / "k.fa"), "--kmer_size", "25", "--DS"])
    b = run([stats, "--reads", str(fa), "--kmers_from_reads", str(fa), "--kmer_size", "25", "--DS"])
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout and len(a.stdout) > 1000


def test_against_reference_binaries_when_present(tmp_path):
    """If the prebuilt reference binaries travelled with the snapshot, diff against them live on a fresh seeded input."""
    ref_stats = os.path.join(orc.REF_DIR, "fastaToKmerCoverageStats")
    ref_r2t = os.path.join(orc.REF_DIR, "ReadsToTranscripts")
    if not (os.path.exists(ref_stats) and os.path.exists(ref_r2t)):
        pytest.skip("oracle/_ref not present")
    import synthdata
    rng = np.random.default_rng(99)
    txs = synthdata.transcriptome(rng, 80, mean_len=800, min_len=200, max_len=3000)
    reads = synthdata.reads_from(rng, txs, 8000, 110, lower_rate=0.05, var_len=True)
    fa = tmp_path / "reads.fa"
    fa.write_bytes(synthdata.fasta_text([">q%d/1 x y" % i for i in range(len(reads))], reads))
    bn, bs = synthdata.bundles_from(rng, txs)
    bf = tmp_path / "bundles.fa"
    bf.write_bytes(synthdata.fasta_text(bn, bs))
    for mode in ("--DS", "--SS"):
        a = run([ref_stats, "--reads", str(fa), "--kmers_from_reads", str(fa), "--num_threads", "1", mode])
        b = run([os.path.join(BIN, "fastaToKmerCoverageStats"), "--reads", str(fa), "--kmers_from_reads", str(fa), mode])
        assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout
    for flags in ([], ["-strand"]):
        oa, ob = tmp_path / "a.out", tmp_path / "b.out"
        common = ["-i", str(fa), "-f", str(bf), "-max_mem_reads", "3000", "-p", "10"] + flags
        assert run([ref_r2t, "-o", str(oa), "-t", "1"] + common).returncode == 0
        assert run([os.path.join(BIN, "ReadsToTranscripts"), "-o", str(ob), "-t", "8"] + common).returncode == 0
        assert oa.read_bytes() == ob.read_bytes() and len(oa.read_bytes()) > 10000
        assert (tmp_path / "a.out.rcts.out").read_bytes() == (tmp_path / "b.out.rcts.out").read_bytes()


def test_dump_sidecar_binary_handoff(tmp_path):
    """`jellyfish dump > kmers.fa` leaves kmers.fa.tgk (packed pairs) next to the FASTA, and
    fastaToKmerCoverageStats --kmers loads it instead of re-parsing the text -- only while it provably describes that
    FASTA (length + content hash).  Output identical either way; any change to the FASTA falls back to the text."""
    jf = os.path.join(BIN, "jellyfish")
    stats = os.path.join(BIN, "fastaToKmerCoverageStats")
    fa = os.path.join(GOLD, "reads.fa")
    db = tmp_path / "m.jf"
    assert run([jf, "count", "-t", "2", "-m", "25", "-s", "1000000", "--canonical", "-o", str(db), fa]).returncode == 0
    kfa = tmp_path / "k.fa"
    with open(kfa, "wb") as out:
        r = subprocess.run([jf, "dump", "-L", "2", str(db)], stdout=out, stderr=subprocess.PIPE, env=ENV, timeout=300)
    assert r.returncode == 0
    side = tmp_path / "k.fa.tgk"
    assert side.exists() and not (tmp_path / "k.fa.tgk.tmp").exists()
    piped = run([jf, "dump", "-L", "2", str(db)])                       # same text through a pipe: no file to sit next to
    assert piped.stdout == kfa.read_bytes()
    nrec = kfa.read_bytes().count(b">")
    assert side.stat().st_size == 40 + 12 * nrec

    cmd = [stats, "--reads", fa, "--kmers", str(kfa), "--kmer_size", "25", "--DS"]
    with_side = run(cmd)
    assert with_side.returncode == 0 and len(with_side.stdout) > 1000
    no_side = subprocess.run(cmd, capture_output=True, env=dict(ENV, TRINITY_GPU_NO_SIDECAR="1"), timeout=300)
    assert no_side.returncode == 0 and no_side.stdout == with_side.stdout
    done = [l for l in with_side.stderr.split(b"\n") if b"done parsing" in l]
    done2 = [l for l in no_side.stderr.split(b"\n") if b"done parsing" in l]
    assert done and done[0].split(b", taking")[0] == done2[0].split(b", taking")[0]      # same "N Kmers, M added"

    # the FASTA edited in place (same length: one count digit changed) -> the hash no longer matches -> text is parsed
    text = bytearray(kfa.read_bytes())
    first_nl = text.index(b"\n")
    assert text[1:first_nl].isdigit()
    text[first_nl - 1] = ord("9") if text[first_nl - 1] != ord("9") else ord("8")
    kfa.write_bytes(bytes(text))
    edited = run(cmd)
    edited_text = subprocess.run(cmd, capture_output=True, env=dict(ENV, TRINITY_GPU_NO_SIDECAR="1"), timeout=300)
    assert edited.returncode == 0 and edited.stdout == edited_text.stdout and edited.stdout != with_side.stdout
    # ... and grown by a record -> length mismatch -> text again
    kfa.write_bytes(bytes(text) + b">7\n" + b"ACGT" * 6 + b"A\n")
    grown = run(cmd)
    grown_text = subprocess.run(cmd, capture_output=True, env=dict(ENV, TRINITY_GPU_NO_SIDECAR="1"), timeout=300)
    assert grown.returncode == 0 and grown.stdout == grown_text.stdout

    # -o writes a sidecar as well; the switch turns the writer off; column format never gets one
    assert run([jf, "dump", "-L", "1", "-o", str(tmp_path / "o.fa"), str(db)]).returncode == 0
    assert (tmp_path / "o.fa.tgk").exists()
    r = subprocess.run([jf, "dump", "-o", str(tmp_path / "n.fa"), str(db)], capture_output=True,
                       env=dict(ENV, TRINITY_GPU_NO_SIDECAR="1"), timeout=300)
    assert r.returncode == 0 and not (tmp_path / "n.fa.tgk").exists()
    assert run([jf, "dump", "-c", "-o", str(tmp_path / "c.txt"), str(db)]).returncode == 0
    assert not (tmp_path / "c.txt.tgk").exists()


def test_jellyfish_count_fastq_equals_fasta(tmp_path):
    """`jellyfish count` on a 4-line FASTQ counts what it counts on the same sequences as FASTA (Trinity hands it FASTA;
    FASTQ is what users run it on by hand)."""
    jf = os.path.join(BIN, "jellyfish")
    entries = orc.read_fasta_inchworm(gold("reads.fa"))
    fa, fq = tmp_path / "r.fa", tmp_path / "r.fq"
    fa.write_text("".join(">%s\n%s\n" % (h, s) for h, _, s in entries))
    fq.write_text("".join("@%s\n%s\n+\n%s\n" % (h, s, "@" * len(s)) for h, _, s in entries))
    outs = []
    for src in (fa, fq):
        db = tmp_path / (src.name + ".jf")
        assert run([jf, "count", "-m", "25", "-s", "1000000", "-C", "-o", str(db), str(src)]).returncode == 0
        d = run([jf, "dump", str(db)])
        assert d.returncode == 0 and len(d.stdout) > 1000
        outs.append(d.stdout)
    assert outs[0] == outs[1]


def test_stats_from_real_jellyfish_dump():
    """fastaToKmerCoverageStats --kmers on REAL jellyfish output (the reference tree's own fixture, see
    tests/golden/make_golden_real_jf.py) against the unmodified reference tool's output, byte for byte."""
    exe = os.path.join(BIN, "fastaToKmerCoverageStats")
    r = run([exe, "--reads", os.path.join(GOLD, "real_jf_reads.fa"), "--kmers", os.path.join(GOLD, "real_jf_dump_head.fa"),
             "--kmer_size", "25", "--DS"])
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == gold("stats_real_jf.expected")
    assert b"done parsing 3000 Kmers, 3000 added" in r.stderr


def _gpu_lists():
    """device lists for TRINITY_GPUS: distinct devices where the box has them, and always the same device twice (two
    contexts on one GPU exercise the same splitting / replication code on a one-GPU box)"""
    import torch
    n = torch.cuda.device_count()
    lists = ["0,0", "0,0,0"]
    if n >= 2:
        lists.append(",".join(str(i) for i in range(min(n, 8))))
    return lists


def test_stats_and_assignment_on_several_gpus(tmp_path):
    """TRINITY_GPUS=0,1,..: the per-read tools replicate the table on every listed device and split every batch of reads
    over them (host/multi_gpu.hpp).  Output must be byte-identical to the single-GPU run == the reference's goldens, in
    all three table modes of the statistics tool (counted from the reads, loaded from a dump, with per-window coverage)
    and for ReadsToTranscripts with small chunks."""
    fa = os.path.join(GOLD, "reads.fa")
    stats = os.path.join(BIN, "fastaToKmerCoverageStats")
    r2t = os.path.join(BIN, "ReadsToTranscripts")
    for gl in _gpu_lists():
        env = dict(ENV, TRINITY_GPUS=gl)
        r = subprocess.run([stats, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "25", "--DS"], capture_output=True,
                           env=env, timeout=300)
        assert r.returncode == 0, (gl, r.stderr.decode()[-2000:])
        assert r.stdout == gold("stats_DS.expected"), gl
        r = subprocess.run([stats, "--reads", fa, "--kmers_from_reads", fa, "--capture_coverage_info"], capture_output=True,
                           env=env, timeout=300)
        assert r.returncode == 0 and r.stdout == gold("stats_capture.expected"), gl
        r = subprocess.run([stats, "--reads", fa, "--kmers", os.path.join(GOLD, "kmers_L2.fa"), "--kmer_size", "25", "--DS"],
                           capture_output=True, env=env, timeout=300)
        assert r.returncode == 0 and r.stdout == gold("stats_kmers_L2.expected"), gl
        out = tmp_path / f"r2c_{gl.replace(',', '_')}.out"
        r = subprocess.run([r2t, "-i", fa, "-f", os.path.join(GOLD, "bundles.fa"), "-o", str(out), "-max_mem_reads", "100",
                            "-p", "10"], capture_output=True, env=env, timeout=300)
        assert r.returncode == 0, (gl, r.stderr.decode()[-2000:])
        assert out.read_bytes() == gold("r2t_p10_chunk100.expected"), gl
        # jellyfish count: chunks dealt round robin to one table per device, tables summed at the end; the database must dump
        # exactly like the single-GPU one
        jf = os.path.join(BIN, "jellyfish")
        one, many = tmp_path / "one.jf", tmp_path / f"many_{gl.replace(',', '_')}.jf"
        assert subprocess.run([jf, "count", "-m", "25", "-s", "1000000", "--canonical", "-o", str(one), fa], capture_output=True,
                              env=ENV, timeout=300).returncode == 0
        r = subprocess.run([jf, "count", "-m", "25", "-s", "1000000", "--canonical", "-o", str(many), fa], capture_output=True,
                           env=dict(env, TRINITY_GPU_COUNT_CHUNK="20000"), timeout=300)
        assert r.returncode == 0, (gl, r.stderr.decode()[-2000:])
        d1 = subprocess.run([jf, "dump", "-L", "1", str(one)], capture_output=True, env=ENV, timeout=300)
        d2 = subprocess.run([jf, "dump", "-L", "1", str(many)], capture_output=True, env=ENV, timeout=300)
        assert d1.returncode == 0 and d2.returncode == 0 and len(d1.stdout) > 1000 and d1.stdout == d2.stdout, gl
        h1 = subprocess.run([jf, "histo", str(one)], capture_output=True, env=ENV, timeout=300)
        h2 = subprocess.run([jf, "histo", str(many)], capture_output=True, env=ENV, timeout=300)
        assert h1.stdout == h2.stdout and len(h1.stdout) > 0, gl
    # a device that does not exist fails loudly (no fallback to fewer GPUs)
    r = subprocess.run([stats, "--reads", fa, "--kmers_from_reads", fa], capture_output=True, env=dict(ENV, TRINITY_GPUS="0,99"),
                       timeout=300)
    assert r.returncode != 0


def test_kmer_size_32(tmp_path):
    """--kmer_size 32 (the reference's maximum, Inchworm/src/KmerCounter.cpp:15-17): stdout byte-identical to the reference
    binary on reads that hold poly-A / poly-T 32-mers -- the all-zero 64-bit key and its reverse complement
    (tests/golden/make_golden_k32.py); jellyfish count -m 32 / dump / histo against the oracle; 33 is refused like the
    reference refuses it."""
    stats = os.path.join(BIN, "fastaToKmerCoverageStats")
    jf = os.path.join(BIN, "jellyfish")
    fa = os.path.join(GOLD, "reads_k32.fa")
    for mode in ("DS", "SS"):
        r = run([stats, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "32", "--" + mode])
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout == gold(f"stats_k32_{mode}.expected")
    r = run([stats, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "32", "--capture_coverage_info"])
    assert r.returncode == 0 and r.stdout == gold("stats_k32_capture.expected")
    r = run([stats, "--reads", fa, "--kmers_from_reads", fa, "--kmer_size", "33"])
    assert r.returncode == 1 and b"exceeds max of 32" in r.stderr
    seqs = []
    for line in gold("reads_k32.fa").split(b"\n"):
        if line.startswith(b">"):
            seqs.append(b"")
        elif seqs:
            seqs[-1] += line
    recs, _ = tg.records_from_sequences(seqs)
    for canonical in (True, False):
        db = tmp_path / f"mer32_{int(canonical)}.jf"
        r = run([jf, "count", "-t", "4", "-m", "32", "-s", "1000000", "-o", str(db)] + (["--canonical"] if canonical else []) + [fa])
        assert r.returncode == 0, r.stderr.decode()
        for L in (1, 2):
            keys, cnts = orc.jf_count(recs, 32, canonical, L)
            assert keys[0] == 0
            expect = "".join(">%d\n%s\n" % (c, tg.packed_to_kmer(k, 32)) for k, c in zip(keys, cnts)).encode()
            d = run([jf, "dump", "-L", str(L), str(db)])
            assert d.returncode == 0 and d.stdout == expect
        keys, cnts = orc.jf_count(recs, 32, canonical, 1)
        bins = orc.jf_histo(cnts)
        h = run([jf, "histo", "-t", "4", "-o", str(tmp_path / "h32.txt"), str(db)])
        expect = "".join("%d %d\n" % (c, bins[c]) for c in range(1, 10002) if bins[c]).encode()
        assert h.returncode == 0 and (tmp_path / "h32.txt").read_bytes() == expect
    # the normalisation pipeline's hand-off at k = 32: dump -> --kmers == counting the reads (no read of exactly 32 bases)
    entries = orc.read_fasta_inchworm(gold("reads_k32.fa"))
    fa2 = tmp_path / "r.fa"
    fa2.write_text("".join(">%s\n%s\n" % (h, s) for h, _, s in entries if len(s) != 32))
    assert run([jf, "count", "-m", "32", "-s", "1000000", "--canonical", "-o", str(tmp_path / "m.jf"), str(fa2)]).returncode == 0
    d = run([jf, "dump", "-L", "1", str(tmp_path / "m.jf")])
    (tmp_path / "k.fa").write_bytes(d.stdout)
    a = run([stats, "--reads", str(fa2), "--kmers", str(tmp_path / "k.fa"), "--kmer_size", "32", "--DS"])
    b = run([stats, "--reads", str(fa2), "--kmers_from_reads", str(fa2), "--kmer_size", "32", "--DS"])
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout and len(a.stdout) > 1000
    assert run([jf, "count", "-m", "33", "-s", "1000", "-o", str(tmp_path / "x.jf"), fa]).returncode != 0
