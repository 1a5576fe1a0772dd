#!/usr/bin/env python
"""Golden vector of the weldmer-counting step: the counts printed by oracle/_ref/weld_count_ref (our 20-line driver around
the UNMODIFIED reference classes NonRedKmerTable / DNAStringStreamFast, compiled from /root/reference by
oracle/Makefile.ref) on the seeded case of tests/weldcase.py.  Run where /root/reference exists:
    python tests/golden/make_golden_weld.py"""
import json, os, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import weldcase

out = {}
for kk in (48, 33, 40):
    reads, cands = weldcase.weld_case(seed=11, kk=kk)
    with tempfile.TemporaryDirectory() as td:
        weldcase.write_fasta(os.path.join(td, "cands.fa"), cands, "c")
        weldcase.write_fasta(os.path.join(td, "reads.fa"), [r for r in reads], "r")
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "weld_count_ref"), str(kk), os.path.join(td, "cands.fa"),
                            os.path.join(td, "reads.fa")], capture_output=True, text=True, check=True)
    out[str(kk)] = [int(x) for x in r.stdout.split()]
    assert len(out[str(kk)]) == len(cands)
json.dump(out, open(os.path.join(HERE, "weld_counts.json"), "w"))
print({k: (len(v), sum(v)) for k, v in out.items()})
