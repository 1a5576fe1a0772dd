"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see the header of oracle/oracle.c).

Also holds the small Python restatements of the reference's three FASTA readers, which the tests use to turn
text into record buffers exactly as each reference tool would see it:
  read_fasta_inchworm   Inchworm/src/Fasta_reader.cpp:82-129 + Fasta_entry.cpp:6-29
  read_fasta_dnastream  Chrysalis/analysis/DNAVector.cc:1456-1501 (DNAStringStreamFast)
  read_bundles          Chrysalis/analysis/DNAVector.cc:856-971 (vecDNAVector::Read) + util/mutil.cc:365-395
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")


def build():
    subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.orc_kc_new.restype = vp; L.orc_kc_new.argtypes = [i32, i32]
        L.orc_kc_free.argtypes = [vp]
        L.orc_kc_size.restype = u64; L.orc_kc_size.argtypes = [vp]
        L.orc_kc_add_kmer_str.argtypes = [vp, C.c_char_p, u32]
        L.orc_kc_add_sequence.argtypes = [vp, C.c_char_p, C.c_int64]
        L.orc_kc_add_records.argtypes = [vp, vp, vp, u64]
        L.orc_cov_stats.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]
        L.orc_jf_count.restype = u64
        L.orc_jf_count.argtypes = [vp, u64, i32, i32, u32, C.POINTER(vp), C.POINTER(vp)]
        L.orc_jf_histo.argtypes = [vp, u64, vp]
        L.orc_free.argtypes = [vp]
        L.orc_rt_new.restype = vp; L.orc_rt_new.argtypes = [i32]
        L.orc_rt_free.argtypes = [vp]
        L.orc_rt_size.restype = u64; L.orc_rt_size.argtypes = [vp]
        L.orc_rt_label.argtypes = [vp, vp, vp, u64, u32]
        L.orc_rt_assign.argtypes = [vp, vp, vp, u64, i32, C.c_float, vp, vp, vp]
        L.orc_entropy.restype = C.c_float; L.orc_entropy.argtypes = [C.c_char_p, i32]
        L.orc_weld_count.argtypes = [vp, u64, i32, vp, vp, u64, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(recs):
    if isinstance(recs, (bytes, bytearray)):
        recs = np.frombuffer(recs, dtype=np.uint8)
    return np.ascontiguousarray(recs, dtype=np.uint8)


class KmerCounter:
    """Inchworm KmerCounter restated (oracle.c: orc_kc_*)."""

    def __init__(self, k=25, ds=True):
        self.k = k
        self._h = C.c_void_p(lib().orc_kc_new(k, int(ds)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_kc_free(self._h)
            self._h = None

    def size(self):
        return lib().orc_kc_size(self._h)

    def add_kmer(self, kmer, count):
        lib().orc_kc_add_kmer_str(self._h, kmer.encode() if isinstance(kmer, str) else kmer, count)

    def add_records(self, recs, offs):
        recs = _u8(recs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lib().orc_kc_add_records(self._h, _p(recs), _p(offs), len(offs) - 1)

    def coverage_stats(self, recs, offs, capture=False):
        recs = _u8(recs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        med = np.zeros(n, np.uint32); mean = np.zeros(n, np.float32); sd = np.zeros(n, np.float32)
        per = np.zeros(int(offs[-1]) if n else 0, np.uint32) if capture else None
        lib().orc_cov_stats(self._h, _p(recs), _p(offs), n, _p(med), _p(mean), _p(sd), _p(per))
        return (med, mean, sd, per) if capture else (med, mean, sd)


def jf_count(recs, k=25, canonical=True, min_count=1):
    """jellyfish count + dump -L restated -> (sorted packed k-mers, counts)."""
    recs = _u8(recs)
    pk, pc = C.c_void_p(), C.c_void_p()
    n = lib().orc_jf_count(_p(recs), recs.nbytes, k, int(canonical), min_count, C.byref(pk), C.byref(pc))
    keys = np.ctypeslib.as_array(C.cast(pk, C.POINTER(C.c_uint64)), shape=(max(n, 1),))[:n].copy()
    cnts = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_uint32)), shape=(max(n, 1),))[:n].copy()
    lib().orc_free(pk); lib().orc_free(pc)
    return keys, cnts


def entropy(window):
    w = window.encode() if isinstance(window, str) else window
    return float(lib().orc_entropy(w, len(w)))


def jf_histo(counts):
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    bins = np.zeros(10002, np.uint64)
    lib().orc_jf_histo(_p(counts), len(counts), _p(bins))
    return bins


class BundleTable:
    """ReadsToTranscripts' labelled NonRedKmerTable restated (oracle.c: orc_rt_*)."""

    def __init__(self, k=25):
        self.k = k
        self._h = C.c_void_p(lib().orc_rt_new(k))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rt_free(self._h)
            self._h = None

    def size(self):
        return lib().orc_rt_size(self._h)

    def label(self, recs, offs, first_index=0):
        recs = _u8(recs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lib().orc_rt_label(self._h, _p(recs), _p(offs), len(offs) - 1, first_index)

    def assign(self, recs, offs, strand=False, min_kmer_entropy=1.5):
        recs = _u8(recs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        best = np.zeros(n, np.int32); pct = np.zeros(n, np.int32); score = np.zeros(n, np.int32)
        lib().orc_rt_assign(self._h, _p(recs), _p(offs), n, int(strand), C.c_float(min_kmer_entropy), _p(best), _p(pct),
                            _p(score))
        return best, pct, score


# ---------------------------------------------------------------------------------------------------------
# the reference's FASTA readers, restated
# ---------------------------------------------------------------------------------------------------------
def _getlines(data):
    """std::getline view of a byte string: list of (line, hit_eof_while_reading)."""
    out = []
    pos = 0
    n = len(data)
    while pos <= n:
        if pos == n:
            out.append((b"", True))      # a getline at EOF: empty line, eof+fail
            break
        nl = data.find(b"\n", pos)
        if nl < 0:
            out.append((data[pos:], True))
            pos = n + 1
            break
        out.append((data[pos:nl], False))
        pos = nl + 1
    return out


def read_fasta_inchworm(data):
    """-> list of (header_without_gt, accession, SEQUENCE).  Whitespace (space, tab, newline) stripped,
    upper-cased; a last record without trailing newline is kept; text before the first '>' is ignored."""
    recs = []
    header = None
    seq = []
    for line in data.split(b"\n"):
        if line[:1] == b">":
            if header is not None:
                recs.append((header, b"".join(seq)))
            header = line
            seq = []
        elif header is not None:
            seq.append(line)
    if header is not None:
        recs.append((header, b"".join(seq)))
    out = []
    for h, s in recs:
        s = s.replace(b" ", b"").replace(b"\t", b"").upper()
        h = h[1:]
        toks = h.replace(b"\t", b" ").split()
        out.append((h.decode(), toks[0].decode() if toks else "", s.decode()))
    return out


def read_fasta_dnastream(data):
    """DNAStringStreamFast::ReadStream/NextToVector -> list of (name_line_with_gt, sequence_verbatim).
    The first sequence line is taken whatever it looks like; a final line that lacks '\\n' is lost
    (so a single-line last record without trailing newline is dropped entirely)."""
    lines = _getlines(data)
    i = 0
    # ReadStream: seek to the first header
    while i < len(lines) and not lines[i][1] and lines[i][0][:1] != b">":
        i += 1
    out = []
    while i < len(lines) and not lines[i][1]:          # m_ifs.good() and m_buf holds a header
        name = lines[i][0]
        i += 1
        if i >= len(lines) or lines[i][1]:             # first sequence line hit EOF: record dropped
            break
        seq = [lines[i][0]]
        i += 1
        while i < len(lines) and not lines[i][1] and lines[i][0][:1] != b">":
            seq.append(lines[i][0])
            i += 1
        out.append((name.decode(), b"".join(seq).decode()))
    return out


def read_bundles(data):
    """vecDNAVector::Read(f,false,false,true,..) -> list of (name, SEQUENCE).  Name = whitespace tokens of the
    header joined by '_'; sequence lines contribute their first token; upper-cased; a last line without
    trailing newline is lost (FlatFileParser::ParseLine returns false once EOF was hit)."""
    out = []
    cur = None
    lines = data.split(b"\n")
    complete = lines[:-1]            # the piece after the last '\n' (possibly empty) never parses
    for line in complete:
        toks = line.replace(b"\t", b" ").split(b" ")
        toks = [t for t in toks if t != b""]
        if not toks:
            continue
        if toks[0][:1] == b">":
            cur = [b"_".join(toks), []]
            out.append(cur)
        elif cur is not None:
            cur[1].append(toks[0])
    return [(n.decode(), b"".join(s).upper().decode()) for n, s in out]


def format_read_name(name):
    """DNAStringStreamFast::formatReadNameString (DNAVector.cc:1504-1514)."""
    while name[:1] == " ":
        name = name[1:]
    name = name.replace(" ", "_")
    while name[-1:] == " ":
        name = name[:-1]
    return name


def weld_count(weldmers, kk, recs, offs):
    """GraphFromFasta weldmer counting restated (oracle.c: orc_weld_count): weldmers = list of kk-character strings;
    -> int32 counts in input order"""
    blob = np.frombuffer(b"".join(w if isinstance(w, bytes) else w.encode() for w in weldmers), dtype=np.uint8)
    assert blob.size == len(weldmers) * kk
    recs = _u8(recs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
    out = np.zeros(len(weldmers), np.int32)
    lib().orc_weld_count(_p(blob) if blob.size else None, len(weldmers), kk, _p(recs), _p(offs), len(offs) - 1, _p(out))
    return out
