"""Seeded synthetic inputs for the parity tests (numpy, CPU).  Shapes follow SURVEY §8(d) at toy sizes."""
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[a] = b


def revcomp(b):
    return COMP[np.frombuffer(b, dtype=np.uint8)][::-1].tobytes()


def transcriptome(rng, ntx, mean_len=1500, min_len=300, max_len=10000):
    lens = np.clip(np.round(rng.lognormal(np.log(mean_len), 0.6, ntx)), min_len, max_len).astype(np.int64)
    return [ALPHA[rng.integers(0, 4, n)].tobytes() for n in lens]


def reads_from(rng, txs, nreads, read_len, err=0.005, n_rate=0.001, weights=None, lower_rate=0.0, var_len=False):
    """single-end style reads (both strands); returns list of bytes"""
    ntx = len(txs)
    if weights is None:
        weights = rng.lognormal(0, 2.0, ntx)
    p = weights / weights.sum()
    which = rng.choice(ntx, size=nreads, p=p)
    out = []
    for t in which:
        tx = txs[t]
        L = read_len if not var_len else int(rng.integers(max(1, read_len // 3), read_len + 1))
        L = min(L, len(tx))
        s = int(rng.integers(0, len(tx) - L + 1))
        r = tx[s:s + L]
        if rng.random() < 0.5:
            r = revcomp(r)
        a = np.frombuffer(r, dtype=np.uint8).copy()
        e = rng.random(L) < err
        if e.any():
            a[e] = ALPHA[(np.searchsorted(ALPHA, a[e]) + rng.integers(1, 4, e.sum())) % 4]
        nmask = rng.random(L) < n_rate
        a[nmask] = ord("N")
        if lower_rate and rng.random() < lower_rate:
            a = np.frombuffer(a.tobytes().lower(), dtype=np.uint8)
        out.append(a.tobytes())
    return out


def bundles_from(rng, txs, max_contigs=4, share_every=7):
    """Inchworm-bundle style records: contigs (transcript pieces) joined by 'X'; every `share_every`-th bundle
    re-uses a piece of an earlier bundle so that some k-mers belong to several bundles (rule R4)."""
    seqs, names = [], []
    i = 0
    b = 0
    pieces_hist = []
    while i < len(txs):
        n = int(rng.integers(1, max_contigs + 1))
        pieces = []
        for tx in txs[i:i + n]:
            a = int(rng.integers(0, max(1, len(tx) // 4)))
            pieces.append(tx[a:a + int(rng.integers(100, max(101, len(tx) - a)))])
        if b % share_every == share_every - 1 and pieces_hist:
            old = pieces_hist[int(rng.integers(0, len(pieces_hist)))]
            pieces.append(old[:min(len(old), 80)])
        pieces_hist.extend(pieces)
        seqs.append(b"X".join(pieces))
        names.append(">s_%d %s" % (b * 3, " ".join(str(int(rng.integers(1, 200))) for _ in pieces)))
        i += n
        b += 1
    return names, seqs


def fasta_text(names, seqs, width=None):
    out = []
    for n, s in zip(names, seqs):
        out.append((n if isinstance(n, bytes) else n.encode()) + b"\n")
        if width:
            for j in range(0, len(s), width):
                out.append(s[j:j + width] + b"\n")
        else:
            out.append(s + b"\n")
    return b"".join(out)
