"""CPU: pins the oracle (oracle/oracle.c + the reader restatements in oracle/oracle_py.py) against outputs of the
unmodified reference binaries committed under tests/golden/ (regenerate: tests/golden/make_golden.py), and -- in the
build container, where /root/reference and oracle/_ref exist -- against the survey's md5 vectors."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import trinityrnaseq_b200 as tg
from oracle import oracle_py as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_DATA = "/root/reference/trinity_ext_sample_data/__regression_tests/test_GraphFromFasta"


def gold(name):
    with open(os.path.join(GOLD, name), "rb") as f:
        return f.read()


def oracle_stats_text(reads_text, ds=True, k=25, capture=False, kmers_text=None):
    """What fastaToKmerCoverageStats prints (single thread), computed by the oracle."""
    entries = orc.read_fasta_inchworm(reads_text)
    kc = orc.KmerCounter(k, ds)
    if kmers_text is None:
        recs, offs = tg.records_from_sequences([s for _, _, s in entries])
        kc.add_records(recs, offs)
    else:
        for h, _, s in orc.read_fasta_inchworm(kmers_text):
            if s == "":
                break
            if len(s) != k:
                continue
            kc.add_kmer(s, int(h.split()[0]) & 0xFFFFFFFF)
    keep = [(a, s) for _, a, s in entries if s != ""]
    recs, offs = tg.records_from_sequences([s for _, s in keep])
    res = kc.coverage_stats(recs, offs, capture=capture)
    lines = ["acc\tmedian_cov\tmean_cov\tstdev\ttid"]
    for i, (acc, s) in enumerate(keep):
        line = tg.format_stats_line(acc, res[0][i], res[1][i], res[2][i])
        if capture:
            n = max(0, len(s) - k + 1)
            line += "\t" + ",".join(str(int(x)) for x in res[3][int(offs[i]):int(offs[i]) + n])
        lines.append(line)
    return ("\n".join(lines) + "\n").encode()


def oracle_r2t_text(reads_text, bundles_text, strand, p=0, max_mem_reads=50000000):
    bundles = orc.read_bundles(bundles_text)
    comp_no = [int("".join(ch for ch in n[3:].split("_")[0] if ch.isdigit() or ch == "-") or 0) for n, _ in bundles]
    brecs, boffs = tg.records_from_sequences([s for _, s in bundles])
    bt = orc.BundleTable(25)
    bt.label(brecs, boffs)
    reads = orc.read_fasta_dnastream(reads_text)
    out = []
    total = 0
    for c0 in range(0, len(reads), max_mem_reads):
        chunk = reads[c0:c0 + max_mem_reads]
        recs, offs = tg.records_from_sequences([s for _, s in chunk])
        best, pct, _ = bt.assign(recs, offs, strand=strand)
        ok = [i for i in range(len(chunk)) if best[i] != -1 and pct[i] >= p]
        ok.sort(key=lambda i: best[i])            # stable: read order inside a bundle
        for i in ok:
            out.append("%d\t%s\t%d%%\t%s\n" % (comp_no[best[i]], orc.format_read_name(chunk[i][0]), pct[i], chunk[i][1]))
        total += len(ok)
    return "".join(out).encode(), total


@pytest.mark.parametrize("tag,ds", [("", True), ("", False), ("_nonl", True), ("_nonl", False)])
def test_stats_from_reads(tag, ds):
    assert oracle_stats_text(gold(f"reads{tag}.fa"), ds=ds) == gold(f"stats{tag}_{'DS' if ds else 'SS'}.expected")


def test_stats_capture_k21_and_dump_loader():
    assert oracle_stats_text(gold("reads.fa"), capture=True) == gold("stats_capture.expected")
    assert oracle_stats_text(gold("reads.fa"), k=21) == gold("stats_k21.expected")
    assert oracle_stats_text(gold("reads.fa"), kmers_text=gold("kmers_L2.fa")) == gold("stats_kmers_L2.expected")


def test_stats_k32():
    """the reference's maximum k (Inchworm/src/KmerCounter.cpp:15-17): poly-A / poly-T 32-mers, the all-zero 64-bit value"""
    for ds in (True, False):
        assert oracle_stats_text(gold("reads_k32.fa"), k=32, ds=ds) == gold(f"stats_k32_{'DS' if ds else 'SS'}.expected")
    assert oracle_stats_text(gold("reads_k32.fa"), k=32, capture=True) == gold("stats_k32_capture.expected")


def test_stats_known_answers():
    """n == 0 -> 0 0 -0 ; n == 1 -> c c -nan ; empty sequence -> no line"""
    txt = gold("stats_DS.expected").decode().splitlines()
    rows = {l.split("\t")[0]: l.split("\t")[1:4] for l in txt[1:]}
    assert rows["short/1"] == ["0", "0", "-0"]
    assert rows["exactK/1"][2] == "-nan" and rows["exactK/1"][0] == rows["exactK/1"][1]
    assert "empty/1" not in rows
    assert rows["name"][0].isdigit()          # accession = header up to the first blank


@pytest.mark.parametrize("tag,mode", [("", "ds"), ("", "strand"), ("_nonl", "ds"), ("_nonl", "strand")])
def test_reads_to_transcripts(tag, mode):
    text, total = oracle_r2t_text(gold(f"reads{tag}.fa"), gold("bundles.fa"), strand=(mode == "strand"))
    assert text == gold(f"r2t{tag}_{mode}.expected")
    assert ("%d\n" % total).encode() == gold(f"r2t{tag}_{mode}.expected.rcts.out")


def test_reads_to_transcripts_pct_and_chunks():
    text, total = oracle_r2t_text(gold("reads.fa"), gold("bundles.fa"), strand=False, p=10, max_mem_reads=100)
    assert text == gold("r2t_p10_chunk100.expected")
    assert ("%d\n" % total).encode() == gold("r2t_p10_chunk100.expected.rcts.out")


def test_r2t_known_answers():
    """SURVEY B3 on the fixture: the last record without '\\n' is dropped by Chrysalis' reader, lower-case reads are
    echoed in their original case, names keep '>' and get '_' for blanks"""
    ds = gold("r2t_ds.expected").decode()
    assert ">last/1" not in gold("r2t_nonl_ds.expected").decode()
    assert gold("r2t_nonl_ds.expected") == gold("r2t_ds.expected")
    assert "\t>name_with_spaces/1_extra_words\t" in ds
    assert any(l.split("\t")[3].islower() for l in ds.splitlines())


def test_jellyfish_restatement_agrees_with_inchworm_counter():
    """J is unpinned against real jellyfish; anchor it on the reference's own counter: for reads longer than k the
    multiset of canonical counts must be identical (any consistent canonicalisation)."""
    entries = orc.read_fasta_inchworm(gold("reads.fa"))
    seqs = [s for _, _, s in entries if len(s) > 25]
    recs, offs = tg.records_from_sequences(seqs)
    for canonical in (True, False):
        keys, cnts = orc.jf_count(recs, 25, canonical, 1)
        kc = orc.KmerCounter(25, canonical)
        kc.add_records(recs, offs)
        assert kc.size() == len(keys)
        kc2 = orc.KmerCounter(25, canonical)
        for kmer, c in zip(keys, cnts):
            kc2.add_kmer(tg.packed_to_kmer(kmer, 25), int(c))
        a = kc.coverage_stats(recs, offs, capture=True)
        b = kc2.coverage_stats(recs, offs, capture=True)
        assert np.array_equal(a[3], b[3])
        if canonical:       # printed representative = lexicographically smaller strand
            for kmer in keys[:200]:
                s = tg.packed_to_kmer(kmer, 25)
                rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
                assert s <= rc
    bins = orc.jf_histo(cnts)
    assert bins.sum() == len(cnts) and bins[0] == 0


needs_ref = pytest.mark.skipif(not (os.path.isdir(REF_DATA) and os.path.exists(os.path.join(orc.REF_DIR, "ReadsToTranscripts"))),
                               reason="needs /root/reference and oracle/_ref (build container only)")


@needs_ref
def test_dump_feeds_reference_stats(tmp_path):
    """the executable cross-check of SURVEY §8c: the REAL reference stats tool loaded from our restated jellyfish dump
    (-L 1) prints exactly what it prints when it counts the reads itself"""
    reads = os.path.join(GOLD, "reads.fa")
    entries = orc.read_fasta_inchworm(gold("reads.fa"))
    recs, _ = tg.records_from_sequences([s for _, _, s in entries if len(s) != 25])   # S3: exactly-k reads are skipped
    keys, cnts = orc.jf_count(recs, 25, True, 1)
    dump = tmp_path / "dump.fa"
    dump.write_text("".join(">%d\n%s\n" % (c, tg.packed_to_kmer(k, 25)) for k, c in zip(keys, cnts)))
    exe = os.path.join(orc.REF_DIR, "fastaToKmerCoverageStats")
    a = subprocess.run([exe, "--reads", reads, "--kmers", str(dump), "--num_threads", "1"], capture_output=True, check=True)
    assert a.stdout == gold("stats_DS.expected")


@needs_ref
def test_survey_md5_vectors(tmp_path):
    """SURVEY Appendix B1/B2 on the reference's own GraphFromFasta fixture, computed by the ORACLE"""
    import gzip
    both = gzip.open(os.path.join(REF_DATA, "both.fa.gz")).read()
    assert hashlib.md5(both).hexdigest() == "2c033f296242d6b4103f5295f619af27"
    contigs = []
    cur = []
    for line in gzip.open(os.path.join(REF_DATA, "inchworm.K25.L25.fa.gz")).read().decode().splitlines():
        if line.startswith(">"):
            if cur:
                contigs.append("".join(cur))
            cur = []
        else:
            cur.append(line)
    contigs.append("".join(cur))
    b = []
    for i in range(0, len(contigs), 3):
        b.append(">s_%d %d\n%s\n" % ((i // 3) * 7, i + 1, "X".join(contigs[i:i + 3])))
    bundles = "".join(b).encode()
    assert hashlib.md5(bundles).hexdigest() == "80f5c57ca9e51ae1c440d17efd518f3f"

    def sorted_md5(text, keys):
        p = tmp_path / "x"
        p.write_bytes(text)
        out = subprocess.run(["sort", "-T", str(tmp_path)] + keys + [str(p)], capture_output=True, check=True,
                             env=dict(os.environ, LC_ALL="C")).stdout
        return hashlib.md5(out).hexdigest()

    text, total = oracle_r2t_text(both, bundles, strand=False, p=10)
    assert total == 54904 and sorted_md5(text, ["-k", "1,1n", "-k3,3nr", "-k2,2"]) == "bd62816cfd364b3a61e870c409ec4f1b"
    text, total = oracle_r2t_text(both, bundles, strand=True, p=10)
    assert total == 54894 and sorted_md5(text, ["-k", "1,1n", "-k3,3nr", "-k2,2"]) == "e49823665730cd83109f927cbc29398c"
    for ds, md5 in ((True, "2bed62c1f583c46973ef2173a1f1d549"), (False, "15434bf43017cc30012661c87fd21be1")):
        st = oracle_stats_text(both, ds=ds).decode().splitlines()[1:]
        cut = "".join("\t".join(l.split("\t")[:4]) + "\n" for l in st).encode()
        assert sorted_md5(cut, []) == md5


def test_real_jellyfish_dump_fixture_pins_format_representative_and_loader():
    """The one piece of REAL jellyfish output in the reference tree (trinity_ext_sample_data/test_Inchworm/
    jellyfish.kmers.fa.gz; first 3000 records committed verbatim by tests/golden/make_golden_real_jf.py):
      * format: `>COUNT\\nKMER\\n`, 25 upper-case bases -- exactly what the J restatement prints;
      * representative: every k-mer is the lexicographically smaller (A<C<G<T) of itself and its reverse complement --
        rule (1) of the restatement, on 3000 of 3000 records (chance: 2^-3000);
      * order: jellyfish's own hash order, NOT sorted -- nothing downstream may rely on ours being sorted;
      * loader: the oracle restatement of fastaToKmerCoverageStats --kmers reproduces the reference binary's output on
        reads whose coverage is decided by those counts."""
    text = gold("real_jf_dump_head.fa")
    lines = text.split(b"\n")
    assert lines[-1] == b"" and len(lines) == 6001
    counts, kmers = lines[0:6000:2], lines[1:6000:2]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for c, km in zip(counts, kmers):
        assert c[:1] == b">" and c[1:].isdigit() and int(c[1:]) >= 1
        assert len(km) == 25 and set(km) <= set(b"ACGT")
        assert km <= km.translate(comp)[::-1]
    assert kmers != sorted(kmers) and len(set(kmers)) == len(kmers)
    # the restatement's own dump text of these (k-mer, count) pairs, in this order, is the file
    ours = b"".join(b">%d\n%s\n" % (int(c[1:]), tg.packed_to_kmer(tg.kmer_to_packed(km.decode()), 25).encode())
                    for c, km in zip(counts, kmers))
    assert ours == text
    assert oracle_stats_text(gold("real_jf_reads.fa"), kmers_text=text) == gold("stats_real_jf.expected")
