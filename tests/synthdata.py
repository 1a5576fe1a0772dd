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


# ---------------------------------------------------------------------------------------------------------
# vectorised generators for the BASELINE-shaped slices (hundreds of thousands to millions of reads in seconds)
# ---------------------------------------------------------------------------------------------------------
def flat_transcriptome(rng, ntx, mean_len=1500, sigma=0.6, min_len=300, max_len=10000):
    """-> (bases uint8[total], offs int64[ntx+1]) ; SURVEY §8d transcriptome shape"""
    lens = np.clip(np.round(rng.lognormal(np.log(mean_len), sigma, ntx)), min_len, max_len).astype(np.int64)
    offs = np.zeros(ntx + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    return ALPHA[rng.integers(0, 4, int(offs[-1]))], offs


def paired_reads_fasta(rng, tx, offs, npairs, read_len, frag_mean=300, frag_sd=30, err=0.005, n_rate=0.001,
                       expr_sigma=2.0, stranded=False):
    """both.fa text of SURVEY §8d: all left reads (`>r<i>/1`), then all right reads (`>r<i>/2`), single-line records.
    left = forward prefix of the fragment, right = reverse complement of its suffix; unstranded libraries flip the
    fragment with probability 1/2.  Substitutions i.i.d. `err`, bases -> N i.i.d. `n_rate`."""
    ntx = len(offs) - 1
    lens = np.diff(offs)
    w = rng.lognormal(0.0, expr_sigma, ntx) * lens
    which = rng.choice(ntx, size=npairs, p=w / w.sum())
    tl = lens[which]
    flen = np.clip(np.round(rng.normal(frag_mean, frag_sd, npairs)).astype(np.int64), read_len, None)
    flen = np.minimum(flen, tl)
    rl = np.minimum(read_len, flen)                      # transcripts shorter than a read give shorter reads
    start = offs[which] + (rng.random(npairs) * (tl - flen + 1)).astype(np.int64)
    ar = np.arange(read_len, dtype=np.int64)
    flip = (rng.random(npairs) < 0.5) & (not stranded)
    comp = COMP

    def take(first, reverse):
        """rows of read_len bases starting at `first` (forward) or ending at first+read_len (reverse complemented)"""
        idx = first[:, None] + (ar[None, :] if not reverse else ar[::-1][None, :])
        m = tx[np.clip(idx, 0, len(tx) - 1)]
        return comp[m] if reverse else m

    # fragment on the forward strand: left = frag[:L], right = rc(frag[-L:]); flipped: left = rc(frag)[:L] = rc(frag[-L:])
    fwd_left = take(start, False)
    fwd_right = take(start + flen - read_len, True)
    left = np.where(flip[:, None], fwd_right, fwd_left)
    right = np.where(flip[:, None], fwd_left, fwd_right)
    out = []
    for mate, m in ((1, left), (2, right)):
        m = m.copy()
        e = rng.random(m.shape) < err
        sub = ALPHA[(np.searchsorted(ALPHA, m[e]) + rng.integers(1, 4, int(e.sum()))) % 4]
        m[e] = sub
        m[rng.random(m.shape) < n_rate] = ord("N")
        names = np.char.add(np.char.add(">r", np.arange(npairs).astype(str)), "/%d\n" % mate)
        rows = [n.encode() + m[i, :rl[i]].tobytes() + b"\n" for i, n in enumerate(names)]
        out.append(b"".join(rows))
    return b"".join(out)


def bundles_fasta(rng, tx, offs, max_per_bundle=25):
    """bundled_iworm_contigs.fasta of SURVEY §8d (C4 shape): contigs = consecutive pieces of every transcript, length ~
    lognormal(ln 500, 0.7) clipped to [100, 20000]; bundles of 1..max_per_bundle contigs joined by 'X', header
    `>s_<comp> <cov>...` (CreateIwormFastaBundle.cc:56-68), component ids ascending with gaps."""
    pieces = []
    for t in range(len(offs) - 1):
        a, e = int(offs[t]), int(offs[t + 1])
        while a < e:
            n = int(np.clip(round(rng.lognormal(np.log(500), 0.7)), 100, 20000))
            if e - a - n < 100:
                n = e - a
            pieces.append((a, a + n))
            a += n
    out, i, comp_id = [], 0, 0
    while i < len(pieces):
        m = int(rng.integers(1, max_per_bundle + 1))
        grp = pieces[i:i + m]
        i += m
        out.append(b">s_%d %s\n" % (comp_id, b" ".join(b"%d" % int(rng.integers(1, 500)) for _ in grp)))
        out.append(b"X".join(tx[a:e].tobytes() for a, e in grp) + b"\n")
        comp_id += int(rng.integers(1, 4))
    return b"".join(out)
