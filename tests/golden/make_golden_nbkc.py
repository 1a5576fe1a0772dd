#!/usr/bin/env python
"""Fixtures for the drop-ins of the two Perl helpers that consume the coverage statistics:
left/right statistics tables (sorted, with unpaired entries, -nan / -0 / huge values) and what the UNMODIFIED reference
scripts print for them (util/support_scripts/nbkc_merge_left_right_stats.pl, nbkc_normalize.pl; perl 5.38 here).
Run in the build container only (needs /root/reference and perl):   python tests/golden/make_golden_nbkc.py
"""
import os
import random
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/util/support_scripts"
CASES = {"a": ["--max_cov", "50", "--min_cov", "1", "--max_CV", "10000"],
         "b": ["--max_cov", "5", "--min_cov", "2", "--max_CV", "1"],
         "c": ["--max_cov=200", "--max_CV=10000"]}


def table(side, n, drop, rnd):
    rows = []
    for i in range(n):
        if (i * 7 + side) % drop == 0:
            continue
        med = rnd.choice([0, 1, 1, 2, 3, 5, 8, 13, 40, 200, 250, 1000, 4000000000])
        mean = med * rnd.uniform(0.8, 1.5) if med else 0
        sd = rnd.choice(["-nan", "-0", "0", "-0.0", "%g" % (mean * rnd.uniform(0, 2))])
        rows.append(("read%05d/%d" % (i, side), str(med), "%g" % mean, sd, "thread:0"))
    rows.sort(key=lambda r: r[0].encode())
    return "acc\tmedian_cov\tmean_cov\tstdev\ttid\n" + "".join("\t".join(r) + "\n" for r in rows)


def main():
    rnd = random.Random(7)
    for side, name, drop in ((1, "nbkc_left.stats", 11), (2, "nbkc_right.stats", 13)):
        open(os.path.join(HERE, name), "w").write(table(side, 1200, drop, rnd))
    left, right = os.path.join(HERE, "nbkc_left.stats"), os.path.join(HERE, "nbkc_right.stats")
    pairs = os.path.join(HERE, "nbkc_pairs.expected")
    with open(pairs, "wb") as f:
        subprocess.run(["perl", REF + "/nbkc_merge_left_right_stats.pl", "--left", left, "--right", right, "--sorted"],
                       check=True, stdout=f, stderr=subprocess.DEVNULL)
    for tag, args in CASES.items():
        r = subprocess.run(["perl", REF + "/nbkc_normalize.pl", "--stats_file", pairs] + args, check=True,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        open(os.path.join(HERE, f"nbkc_selected_{tag}.expected"), "wb").write(r.stdout)
        open(os.path.join(HERE, f"nbkc_selected_{tag}.stderr"), "wb").write(r.stderr)
    r = subprocess.run(["perl", REF + "/nbkc_normalize.pl", "--stats_file", left, "--max_cov", "30", "--min_cov", "1",
                        "--max_CV", "100"], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    open(os.path.join(HERE, "nbkc_selected_single.expected"), "wb").write(r.stdout)
    print("written")


if __name__ == "__main__":
    main()
