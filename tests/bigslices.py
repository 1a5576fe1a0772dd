"""BASELINE-shaped parity slices shared by tests/golden/make_golden_big.py (which runs the unmodified reference binaries
on them in the build container) and tests/test_gpu_bigslice.py (which runs the GPU executables on the GPU box)."""
import gzip
import hashlib
import os
import subprocess

import numpy as np

import synthdata

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV = dict(os.environ, LC_ALL="C")

# name -> what to generate / which tool / which modes.  "env" is applied to the GPU executable only.
SLICES = {
    # configs[1] shape: PE 2x100, lognormal expression, 0.5 % substitutions, 0.1 % N
    "c2_2x100_1M": dict(kind="synthetic", tool="stats", seed=20251017, ntx=4000, npairs=500_000, read_len=100,
                        modes=["DS", "SS"], env={}),
    # configs[2] shape: PE 2x150; 1 MB partitions force the table beyond 512 partitions, i.e. the coarse-log + refine path
    "c3_2x150_700k": dict(kind="synthetic", tool="stats", seed=20251018, ntx=3000, npairs=350_000, read_len=150,
                          modes=["DS"], env={"TG_PART_MB": "1", "TG_COUNT_MODE": "log"}),
    # configs[3] shape: reads against bundles cut from the same transcriptome
    "c4_r2t_300k": dict(kind="synthetic", tool="r2t", seed=20251019, ntx=3000, npairs=150_000, read_len=100,
                        modes=["ds", "strand"], env={}),
    # the reference's own fixtures: md5s there are the contract
    "ref_graphfromfasta_stats": dict(kind="ref_gff", tool="stats", modes=["DS", "SS"], env={}),
    "ref_graphfromfasta_r2t": dict(kind="ref_gff", tool="r2t", modes=["ds", "strand"], env={}),
    # (the survey's md5 for this one is `cut -f1-4 | sort` WITH the header line)
    "ref_sample_stats": dict(kind="ref_sample", tool="stats", modes=["DS"], env={}, keep_header=True),
}

# SURVEY.md Appendix B (generated from the reference during the survey)
SURVEY_MD5 = {
    ("ref_graphfromfasta_r2t", "ds"): "bd62816cfd364b3a61e870c409ec4f1b",
    ("ref_graphfromfasta_r2t", "strand"): "e49823665730cd83109f927cbc29398c",
    ("ref_graphfromfasta_stats", "DS"): "2bed62c1f583c46973ef2173a1f1d549",
    ("ref_graphfromfasta_stats", "SS"): "15434bf43017cc30012661c87fd21be1",
    ("ref_sample_stats", "DS"): "bf0dbda13fa91a6c03b16c3316e78f71",
}


def gff_bundles(contigs_fa_text):
    """2,025 bundles of 3 contigs, ids 0,7,14,..., header `>s_<id> <first line number>`"""
    contigs, cur = [], []
    for line in contigs_fa_text.decode().splitlines():
        if line.startswith(">"):
            if cur:
                contigs.append("".join(cur))
            cur = []
        else:
            cur.append(line)
    contigs.append("".join(cur))
    b = [">s_%d %d\n%s\n" % ((i // 3) * 7, i + 1, "X".join(contigs[i:i + 3])) for i in range(0, len(contigs), 3)]
    return "".join(b).encode()


def materialise(name, td):
    """write the slice's input files into td -> {"reads": path[, "bundles": path]}"""
    spec = SLICES[name]
    files = {"reads": os.path.join(td, name + ".reads.fa")}
    if spec["kind"] == "synthetic":
        rng = np.random.default_rng(spec["seed"])
        tx, offs = synthdata.flat_transcriptome(rng, spec["ntx"])
        with open(files["reads"], "wb") as f:
            f.write(synthdata.paired_reads_fasta(rng, tx, offs, spec["npairs"], spec["read_len"]))
        if spec["tool"] == "r2t":
            files["bundles"] = os.path.join(td, name + ".bundles.fa")
            with open(files["bundles"], "wb") as f:
                f.write(synthdata.bundles_fasta(rng, tx, offs))
    elif spec["kind"] == "ref_gff":
        with open(files["reads"], "wb") as f:
            f.write(gzip.open(os.path.join(GOLD, "ref", "both.fa.gz")).read())
        if spec["tool"] == "r2t":
            files["bundles"] = os.path.join(td, name + ".bundles.fa")
            with open(files["bundles"], "wb") as f:
                f.write(gff_bundles(gzip.open(os.path.join(GOLD, "ref", "inchworm.K25.L25.fa.gz")).read()))
    elif spec["kind"] == "ref_sample":
        with open(files["reads"], "wb") as f:
            f.write(gzip.open(os.path.join(GOLD, "ref", "reads.left.fa.gz")).read())
            f.write(gzip.open(os.path.join(GOLD, "ref", "reads.right.fa.gz")).read())
    return files


def _sorted_md5(data, keys, td):
    p = os.path.join(td, "sort.in")
    with open(p, "wb") as f:
        f.write(data)
    out = subprocess.run(["sort", "-T", td] + keys + [p], capture_output=True, check=True, env=ENV).stdout
    return hashlib.md5(out).hexdigest()


def stats_md5(stdout, td, keep_header=False):
    """`tail -n +2 | cut -f1-4 | sort | md5sum`: what downstream reads of a stats file"""
    lines = stdout.split(b"\n")[0 if keep_header else 1:]
    cut = b"".join(b"\t".join(l.split(b"\t")[:4]) + b"\n" for l in lines if l)
    return _sorted_md5(cut, [], td)


def r2t_md5(path, td):
    """`sort -k 1,1n -k3,3nr -k2,2 | md5sum` (Trinity:2258)"""
    with open(path, "rb") as f:
        return _sorted_md5(f.read(), ["-k", "1,1n", "-k3,3nr", "-k2,2"], td)
