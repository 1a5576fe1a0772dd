#!/usr/bin/env python
"""SASS evidence of the built library: per kernel the instruction count, the most frequent mnemonics and the instructions
that show how it talks to memory and to the warp.  Needs only cuobjdump (no GPU).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (--full FILE.gz also writes the hot kernels' listing)"""
import argparse, collections, gzip, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "trinityrnaseq_b200", "lib", "libtrinity_gpu.so")
HOT = ("k_log_tiles", "k_log_replay", "k_cov_stats", "k_assign", "k_locus_tiles")
EVID = re.compile(r"^(UBLKCP|UBLKPF|SYNCS|LDG\.E\.ENL2\.256|REDUX|CREDUX|VOTE|VOTEU|SHFL|BREV|REDG|ATOMG|ATOMS|STL|LDL|BAR|MATCH|HMMA|UTC|TCGEN)")
ap = argparse.ArgumentParser()
ap.add_argument("--full", default="")
a = ap.parse_args()
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur, listing = collections.OrderedDict(), None, {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)).split("(")[0]
        cur = re.sub(r"^void ", "", cur)
        kernels.setdefault(cur, []); listing.setdefault(cur, [])
        continue
    if cur is None:
        continue
    listing[cur].append(line)
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        kernels[cur].append(m.group(1))
print("SASS of trinityrnaseq_b200/lib/libtrinity_gpu.so (cuobjdump -sass, sm_100a; tools/sass_summary.py), per kernel: instruction\n"
      "count, the 14 most frequent mnemonics, and the instructions that show HOW the kernel talks to memory / the warp (TMA bulk\n"
      "copies = UBLKCP, mbarrier = SYNCS.*, 256-bit loads = LDG.E.ENL2.256, warp reductions = REDUX/CREDUX, votes, shuffles, bit\n"
      "reversal, atomics, local-memory spills = STL/LDL).  No tensor-core instruction anywhere (nothing here is a contraction).\n")
for k in sorted(kernels):
    ins = kernels[k]
    if not ins or not k.startswith("tg::"):
        continue
    top = collections.Counter(ins).most_common(14)
    ev = collections.Counter(i for i in ins if EVID.match(i))
    print(f"== {k}: {len(ins)} instructions")
    print("   top: " + ", ".join(f"{n} {c}" for n, c in top))
    print("   evidence: " + ", ".join(f"{n} x{c}" for n, c in sorted(ev.items())) + "\n")
if a.full:
    with gzip.open(a.full, "wt") as f:
        for k in sorted(listing):
            if any(h in k for h in HOT):
                f.write(f"==== {k}\n" + "\n".join(listing[k]) + "\n")
