"""Random-access roofline sweep: GUPS probe (tg_gups) over table sizes from L2-resident to tens of GiB."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trinityrnaseq_b200 as tg

ctx = tg.Context(0)
rows = []
for mib in (32, 64, 96, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536):
    slots = mib * (1 << 20) // 16
    nops = 1 << 29
    r = {"table_mib": mib}
    for mode, name in ((0, "load16"), (1, "load8_red"), (2, "cas_red")):
        ms = ctx.gups(slots, nops, mode, reps=2)
        r[name + "_gops"] = round(nops / ms / 1e6, 2)
    rows.append(r)
    print(json.dumps(r), flush=True)
