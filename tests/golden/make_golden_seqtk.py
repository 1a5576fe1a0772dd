#!/usr/bin/env python
"""Goldens for the seqtk-trinity drop-in: small FASTQ / FASTA inputs with every read-name format and
every malformation the reference distinguishes, and what the UNMODIFIED reference tool (oracle/_ref/seqtk-trinity, built by
oracle/Makefile.ref from trinity-plugins/seqtk-trinity/seqtk.c) prints and returns for them.
Run in the build container only:   python tests/golden/make_golden_seqtk.py
"""
import json
import os
import subprocess

import numpy as np

GOLD = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(GOLD, "seqtk")
ROOT = os.path.dirname(os.path.dirname(GOLD))
REF = os.path.join(ROOT, "oracle", "_ref", "seqtk-trinity")


def fastq(names, rng, n=40, lo=30, hi=160, crlf=False, multiline=False):
    out = []
    eol = "\r\n" if crlf else "\n"
    for i in range(n):
        L = int(rng.integers(lo, hi))
        s = "".join(rng.choice(list("ACGTNacgtRY"), L, p=[.22, .22, .22, .22, .03, .02, .02, .02, .01, .01, .01]))
        q = "".join(chr(int(c)) for c in rng.integers(35, 74, L))
        if multiline and i % 3 == 0 and L > 60:
            s = s[:50] + eol + s[50:]
            q = q[:40] + eol + q[40:]
        out.append("@%s%s%s%s+%s%s%s" % (names(i), eol, s, eol, eol, q, eol))
    return "".join(out)


def cases():
    rng = np.random.default_rng(8)
    c = {}
    c["old_1.fq"] = fastq(lambda i: "61DFRAAXX100204:1:100:%d:%d/1" % (10000 + i, 3000 + i), rng)
    c["old_2.fq"] = fastq(lambda i: "61DFRAAXX100204:1:100:%d:%d/2" % (10000 + i, 3000 + i), rng)
    c["new_1.fq"] = fastq(lambda i: "M01581:927:000000000-ARTAL:1:1101:%d:%d 1:N:0:1" % (19000 + i, 2000 + i), rng)
    c["new_2.fq"] = fastq(lambda i: "M01581:927:000000000-ARTAL:1:1101:%d:%d 2:N:0:1" % (19000 + i, 2000 + i), rng)
    c["bare.fq"] = fastq(lambda i: "SRR1.%d" % i, rng)
    c["bare_comment.fq"] = fastq(lambda i: "SRR1.%d length=76 x" % i, rng)
    c["fwdrev_1.fq"] = fastq(lambda i: "r%d_forward/1" % i if i % 2 else "r%d_reverse/2 c" % i, rng)
    c["crlf.fq"] = fastq(lambda i: "c%d/1" % i, rng, crlf=True)
    c["multiline.fq"] = fastq(lambda i: "m%d 1:N:0" % i, rng, multiline=True)
    c["tabs.fq"] = fastq(lambda i: "t%d\t1:Y:0 rest" % i, rng, n=10)
    c["mixed_formats.fq"] = fastq(lambda i: ["a%d/1", "b%d 1:N:0:2", "c%d", "d%d_forward"][i % 4] % i, rng)
    c["plain.fa"] = "".join(">f%d desc\n%s\n" % (i, "".join(rng.choice(list("ACGT"), 70))) for i in range(20))
    c["wrapped.fa"] = "junk before\n" + "".join(">w%d/1\n%s\n%s\n\n" % (i, "ACGTTGCA" * 5, "GGCC" * 3) for i in range(10))
    c["qual_starts_with_at.fq"] = "@q1/1\nACGTACGTAC\n+\n@IIIIIIIII\n@q2/1\nTTTTGGGGCC\n+q2/1\n+IIIIIIIII\n"
    c["seq_shorter_than_qual.fq"] = "@u1/1\nACGTACGT\n+\nIIIIIIIIII\n@u2/1\nACGT\n+\nIIII\n"
    c["qual_shorter_than_seq.fq"] = "@u1/1\nACGTACGTAC\n+\nIIIIII\n"
    c["truncated_no_qual.fq"] = "@u1/1\nACGTACGTAC\n+"
    c["empty_seq.fq"] = "@e0/1\nACGT\n+\nIIII\n@e1/1\n\n+\n\n@e2/1\nACGT\n+\nIIII\n"
    c["empty_seq.fa"] = ">x1/1\nACGT\n>x2/1\n>x3/1\nAC\n"
    c["no_records.txt"] = "just text\nwithout any header\n"
    c["empty.txt"] = ""
    c["header_only.fq"] = "@"
    c["no_trailing_newline.fq"] = "@n1/1\nACGT\n+\nIIII\n@n2/1\nGGCC\n+\nIIII"
    c["one_char_names.fa"] = ">a\nACGT\n>/\nACGT\n>1\nACGT\n"
    return c


RUNS = [  # (input, argv after `seq`)
    ("old_1.fq", ["-A", "-R", "1"]), ("old_1.fq", ["-A", "-R", "2"]), ("old_2.fq", ["-A", "-R", "2"]), ("old_2.fq", ["-A", "-R", "1"]),
    ("new_1.fq", ["-A", "-R", "1"]), ("new_1.fq", ["-A", "-R", "2"]), ("new_2.fq", ["-A", "-R", "2"]), ("new_2.fq", ["-A", "-r", "-R", "2"]),
    ("new_1.fq", ["-A", "-C", "-R", "2"]), ("bare.fq", ["-A", "-R", "1"]), ("bare.fq", ["-A", "-R", "2", "-r"]), ("bare_comment.fq", ["-A", "-R", "2"]),
    ("fwdrev_1.fq", ["-A", "-R", "1"]), ("crlf.fq", ["-A", "-R", "1"]), ("multiline.fq", ["-A", "-R", "1"]), ("tabs.fq", ["-A", "-R", "1"]),
    ("mixed_formats.fq", ["-A", "-R", "1"]), ("plain.fa", ["-A", "-R", "1"]), ("plain.fa", ["-R", "2", "-l", "30"]), ("wrapped.fa", ["-A", "-R", "1"]),
    ("old_1.fq", ["-R", "1"]), ("old_1.fq", ["-R", "1", "-r", "-l", "25"]), ("old_1.fq", ["-A", "-R", "1", "-U"]), ("old_1.fq", ["-A", "-R", "1", "-N"]),
    ("old_1.fq", ["-A", "-R", "1", "-L", "100"]), ("old_1.fq", ["-A", "-R", "1", "-1"]), ("old_1.fq", ["-A", "-R", "1", "-2"]),
    ("old_1.fq", ["-A", "-R", "1", "-q", "20"]), ("old_1.fq", ["-A", "-R", "1", "-q", "20", "-n", "N"]), ("old_1.fq", ["-A", "-R", "1", "-q", "10", "-X", "30", "-Q", "35"]),
    ("qual_starts_with_at.fq", ["-A", "-R", "1"]), ("seq_shorter_than_qual.fq", ["-A", "-R", "1"]), ("qual_shorter_than_seq.fq", ["-A", "-R", "1"]),
    ("truncated_no_qual.fq", ["-A", "-R", "1"]), ("empty_seq.fq", ["-A", "-R", "1"]), ("empty_seq.fa", ["-A", "-R", "1"]),
    ("no_records.txt", ["-A", "-R", "1"]), ("empty.txt", ["-A", "-R", "1"]), ("header_only.fq", ["-A", "-R", "1"]),
    ("no_trailing_newline.fq", ["-A", "-R", "1"]), ("one_char_names.fa", ["-A", "-R", "1"]), ("old_1.fq", ["-A"]), ("old_1.fq", ["-A", "-R", "3"]),
]


def main():
    os.makedirs(HERE, exist_ok=True)
    for name, text in cases().items():
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(text.encode())
    ref = os.path.abspath(REF)
    expect = []
    for i, (inp, args) in enumerate(RUNS):
        for via_stdin in (False, True):
            path = os.path.join(HERE, inp)
            if via_stdin:
                with open(path, "rb") as f:
                    r = subprocess.run([ref, "seq"] + args + ["-"], stdin=f, capture_output=True)
            else:
                r = subprocess.run([ref, "seq"] + args + [path], capture_output=True)
            out_name = "run%02d%s.out" % (i, "_stdin" if via_stdin else "")
            with open(os.path.join(HERE, out_name), "wb") as f:
                f.write(r.stdout)
            expect.append({"input": inp, "args": args, "stdin": via_stdin, "rc": r.returncode, "stdout": out_name,
                           "stderr": r.stderr.decode(errors="replace").replace(path, "<PATH>")})
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(expect, f, indent=1)
    print(len(expect), "runs recorded in", HERE)


if __name__ == "__main__":
    main()
