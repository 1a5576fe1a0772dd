"""Host-side mirror of the reference's in-process interfaces for this path, on top of the C ABI.

  KmerCounter      <- Inchworm/src/KmerCounter.{hpp,cpp} (add_sequence / add_kmer / get counts) plus the
                      per-read statistics of Inchworm/src/fastaToKmerCoverageStats.cpp and the
                      jellyfish count/dump/histo trio (Trinity:2612-2632)
  BundleKmerTable  <- Chrysalis/analysis/NonRedKmerTable.{h,cc} as used by ReadsToTranscripts.cc

Every method forwards to libtrinity_gpu; nothing is computed in Python.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def kmer_to_packed(kmer):
    v = 0
    for ch in kmer.upper():
        v = (v << 2) | _CODE[ch]
    return v


def packed_to_kmer(v, k):
    return "".join("ACGT"[(int(v) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def records_from_sequences(seqs):
    """list of str/bytes -> (record buffer, offs): each sequence followed by one '\\n' terminator."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) + 1 for b in bs], dtype=np.uint64)
    recs = np.frombuffer(b"".join(b + b"\n" for b in bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return recs, offs


def format_stats_line(acc, median, mean, stdev, tid=0):
    """One output line of fastaToKmerCoverageStats (fastaToKmerCoverageStats.cpp:140-148): iostream default
    float formatting == printf %g, and the x86 default NaN prints as -nan."""
    def g(x):
        x = np.float32(x)
        if np.isnan(x):
            return "-nan" if np.signbit(x) else "nan"
        return "%g" % float(x)
    return f"{acc}\t{int(median)}\t{g(mean)}\t{g(stdev)}\tthread:{tid}"


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _addr(p):
    return p.value if isinstance(p, C.c_void_p) else int(p)


def _as_u8(recs):
    if isinstance(recs, (bytes, bytearray)):
        recs = np.frombuffer(recs, dtype=np.uint8)
    return np.ascontiguousarray(recs, dtype=np.uint8)


class Context:
    """One GPU (tg_ctx).  Fails with TrinityGpuError(TG_ERR_NOGPU) when no B200 is visible."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(_lib.lib().tg_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            _lib.lib().tg_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self):
        sm, fr, tot = C.c_int(), C.c_uint64(), C.c_uint64()
        check(_lib.lib().tg_device_info(self._h, C.byref(sm), C.byref(fr), C.byref(tot)))
        return {"sm_count": sm.value, "free_bytes": fr.value, "total_bytes": tot.value}

    def sync(self):
        check(_lib.lib().tg_sync(self._h))

    def log_overflow_check(self):
        """sync; True when a table-less log append overflowed a bin since the last check (flag cleared)"""
        f = C.c_int(0)
        check(_lib.lib().tg_log_overflow_check(self._h, C.byref(f)))
        return bool(f.value)

    def valid_windows_dev(self, d_recs, nbytes, k):
        """number of k-mer windows without a non-ACGT byte in a device record buffer (independent of the count kernels)"""
        n = C.c_uint64(0)
        check(_lib.lib().tg_valid_windows_dev(self._h, d_recs, nbytes, k, C.byref(n)))
        return int(n.value)

    def records_hold(self, recs):
        """declare `recs` (a uint8 array, e.g. from pinned()) immutable until records_release: ONE upload shared by the
        count / statistics / assignment calls that are given this same array"""
        a = _as_u8(recs)
        check(_lib.lib().tg_records_hold(self._h, _ptr(a), a.nbytes))
        self._held = a

    def records_release(self):
        check(_lib.lib().tg_records_release(self._h))
        self._held = None

    def records_pin_dev(self, d_recs=None, d_offs=None, nreads=0):
        """declare a device record buffer immutable (its locus order is computed once); no arguments: unpin"""
        check(_lib.lib().tg_records_pin_dev(self._h, d_recs, d_offs, nreads))

    def locus_prepare_dev(self, k, recompute=False):
        """queue the locus order of the pinned buffer on the second stream (it overlaps the work queued next)"""
        check(_lib.lib().tg_locus_prepare_dev(self._h, int(k), int(bool(recompute))))

    def launch_count(self):
        return int(_lib.lib().tg_launch_count(self._h))

    def set(self, key, value):
        """tuning knob (tg_ctx_set): count_mode, batch_bytes, part_bytes, log_bytes, replay_prefetch, ..."""
        check(_lib.lib().tg_ctx_set(self._h, str(key).encode(), str(value).encode()))

    def kernel_times(self):
        """{kernel name: (total ms, launches)} since the last call; needs set("kernel_timing", 1)"""
        buf = C.create_string_buffer(1 << 14)
        check(_lib.lib().tg_kernel_times(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split("\t")
            out[name] = (float(ms), int(n))
        return out

    # -- raw device memory (bench / tests) ------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        check(_lib.lib().tg_dev_alloc(self._h, nbytes, C.byref(p)))
        return p

    def dev_records_alloc(self, nbytes):
        p = C.c_void_p()
        check(_lib.lib().tg_dev_records_alloc(self._h, nbytes, C.byref(p)))
        return p

    def dev_free(self, p):
        check(_lib.lib().tg_dev_free(self._h, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        check(_lib.lib().tg_memcpy_h2d(self._h, dptr, _ptr(arr), arr.nbytes))

    def d2h(self, dptr, nbytes_or_arr, dtype=np.uint8, offset=0):
        if isinstance(nbytes_or_arr, np.ndarray):
            out = nbytes_or_arr
        else:
            out = np.empty(nbytes_or_arr // np.dtype(dtype).itemsize, dtype=dtype)
        src = C.c_void_p(dptr.value + offset)
        check(_lib.lib().tg_memcpy_d2h(self._h, _ptr(out), src, out.nbytes))
        return out

    def d2d(self, dst, src, nbytes, dst_off=0, src_off=0):
        check(_lib.lib().tg_memcpy_d2d(self._h, C.c_void_p(_addr(dst) + dst_off), C.c_void_p(_addr(src) + src_off), nbytes))

    def memset(self, dst, value, nbytes):
        check(_lib.lib().tg_memset_dev(self._h, dst, value, nbytes))

    def pinned(self, shape, dtype):
        """numpy array backed by pinned host memory (tg_host_alloc); keep the returned owner alive."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _lib.lib().tg_host_alloc(max(n, 1))
        if not p:
            raise MemoryError("tg_host_alloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr, _PinnedOwner(p)

    def timer_start(self):
        check(_lib.lib().tg_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(_lib.lib().tg_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def gups(self, slots, nops, mode, reps=3):
        ms = C.c_float()
        check(_lib.lib().tg_gups(self._h, slots, nops, mode, reps, C.byref(ms)))
        return ms.value

    def synth_reads_dev(self, tx, tx_offs, tx_cum, npairs, read_len, frag_mean=300, frag_sd=30,
                        err_per_million=5000, n_per_million=1000, seed=20251017, stranded=False):
        """Generate 2*npairs reads on the device; returns the device record buffer pointer and its size."""
        tx = _as_u8(tx)
        tx_offs = np.ascontiguousarray(tx_offs, dtype=np.uint64)
        tx_cum = np.ascontiguousarray(tx_cum, dtype=np.uint64)
        nbytes = 2 * npairs * (read_len + 1)
        d = self.dev_records_alloc(nbytes)
        check(_lib.lib().tg_synth_reads_dev(self._h, _ptr(tx), _ptr(tx_offs), _ptr(tx_cum), len(tx_cum), npairs,
                                            read_len, frag_mean, frag_sd, err_per_million, n_per_million, seed,
                                            int(stranded), d))
        return d, nbytes


class _PinnedOwner:
    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            if self.p:
                _lib.lib().tg_host_free(self.p)
                self.p = None
        except Exception:
            pass


class _Table:
    KIND = _lib.TG_TABLE_COUNT

    def __init__(self, ctx, k=25, expected_keys=1 << 20, geometry=None):
        self.ctx, self.k = ctx, k
        self._h = C.c_void_p()
        if geometry is None:
            check(_lib.lib().tg_table_create(ctx._h, self.KIND, k, expected_keys, C.byref(self._h)))
        else:
            subcap, nparts, part0, nlocal = geometry
            check(_lib.lib().tg_table_create_sharded(ctx._h, self.KIND, k, subcap, nparts, part0, nlocal,
                                                     C.byref(self._h)))

    def geometry(self):
        """(slots_per_partition, nparts, part0, nlocal) -- see tg_table_create_sharded"""
        sc, n, p0, nl = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(_lib.lib().tg_table_geometry(self._h, C.byref(sc), C.byref(n), C.byref(p0), C.byref(nl)))
        return sc.value, n.value, p0.value, nl.value

    def resize(self, slots_per_partition):
        check(_lib.lib().tg_table_resize(self._h, slots_per_partition))

    def slots_dev(self):
        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().tg_table_slots_dev(self._h, C.byref(p), C.byref(n)))
        return p, n.value

    def set_distinct(self, n):
        check(_lib.lib().tg_table_set_distinct(self._h, int(n)))

    def close(self):
        if self._h:
            _lib.lib().tg_table_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self):
        cap, n = C.c_uint64(), C.c_uint64()
        check(_lib.lib().tg_table_info(self._h, C.byref(cap), C.byref(n)))
        return {"capacity": cap.value, "distinct": n.value}

    def size(self):
        return self.info()["distinct"]

    def clear(self):
        check(_lib.lib().tg_table_clear(self._h))

    def reserve(self, additional):
        check(_lib.lib().tg_table_reserve(self._h, additional))


class KmerCounter(_Table):
    """k-mer -> count table.  is_ds mirrors KmerCounter(kmer_length, is_ds) (KmerCounter.cpp:13-22)."""
    KIND = _lib.TG_TABLE_COUNT

    def __init__(self, ctx, k=25, is_ds=True, expected_keys=1 << 20, geometry=None):
        super().__init__(ctx, k, expected_keys, geometry)
        self.is_ds = bool(is_ds)

    @classmethod
    def sharded(cls, ctx, k, is_ds, slots_per_partition, nparts, part0, nlocal):
        """the shard holding partitions [part0, part0+nlocal) of a table of nparts partitions"""
        return cls(ctx, k, is_ds, geometry=(slots_per_partition, nparts, part0, nlocal))

    def count_min(self, min_count):
        """number of k-mers with count >= min_count (what `jellyfish dump -L min_count` would print)"""
        n = C.c_uint64()
        check(_lib.lib().tg_table_count_min(self._h, min_count, C.byref(n)))
        return n.value

    def compact_into(self, min_count, dst):
        """copy the k-mers with count >= min_count into dst (`jellyfish dump -L` without leaving the device)"""
        check(_lib.lib().tg_table_compact_into(self._h, min_count, dst._h))

    def compacted(self, min_count, load=0.45):
        """new table holding only the k-mers with count >= min_count, same partition range"""
        n = self.count_min(min_count)
        subcap, nparts, part0, nlocal = self.geometry()
        sub = max(int(n / load / nlocal) + 64, 64)
        dst = KmerCounter.sharded(self.ctx, self.k, self.is_ds, sub, nparts, part0, nlocal)
        self.compact_into(min_count, dst)
        return dst

    def partition_dev(self, d_recs, nbytes, nbins, cap, d_keys, d_cursor, d_hpoly, canonical=None):
        """phase 1 of the sharded count: k-mer occurrences -> caller-owned log bins (tg_count_partition_dev)"""
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_count_partition_dev(self.ctx._h, d_recs, nbytes, self.k, int(can), nbins, cap, d_keys,
                                                d_cursor, d_hpoly))

    def partition_peers_dev(self, d_recs, nbytes, nbins, cap, my_rank, owner_keys, d_cursor, d_hpoly, canonical=None):
        """phase 1 fused with the exchange: entries stored into segment [my_rank] of the owners' receive logs
        (tg_count_partition_peers_dev); owner_keys = device pointers of every rank's log"""
        can = self.is_ds if canonical is None else canonical
        ptrs = (C.c_void_p * len(owner_keys))(*[_addr(p) for p in owner_keys])
        check(_lib.lib().tg_count_partition_peers_dev(self.ctx._h, d_recs, nbytes, self.k, int(can), nbins, cap,
                                                      len(owner_keys), my_rank, ptrs, d_cursor, d_hpoly))

    def replay_log_dev(self, d_keys, d_cursor, d_hpoly, nsrc, cap):
        """phase 2: received log [nsrc][nlocal][cap] -> this shard (tg_table_replay_log_dev)"""
        check(_lib.lib().tg_table_replay_log_dev(self._h, d_keys, d_cursor, d_hpoly, nsrc, cap))

    def add_records(self, recs, canonical=None):
        """jellyfish count / KmerCounter::add_sequence over a record buffer."""
        recs = _as_u8(recs)
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_count_reads(self._h, _ptr(recs), recs.nbytes, int(can)))

    def add_records_dev(self, d_recs, nbytes, canonical=None):
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_count_reads_dev(self._h, d_recs, nbytes, int(can)))

    def set_count_floor(self, min_count):
        """statistics see the table as `jellyfish dump -L min_count` would leave it (a view; nothing is rebuilt)"""
        check(_lib.lib().tg_table_set_count_floor(self._h, int(min_count)))

    def add_read_records_dev(self, d_recs, d_offs, nreads, canonical=None):
        """count read by read in locus order (tg_count_records_dev): needs the read offsets, no k-mer log"""
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_count_records_dev(self._h, d_recs, d_offs, nreads, int(can)))

    def add_kmers(self, packed_keys, counts, canonical=None):
        """KmerCounter::add_kmer(kmer, count) for many k-mers (the `--kmers` dump loader)."""
        keys = np.ascontiguousarray(packed_keys, dtype=np.uint64)
        vals = np.ascontiguousarray(counts, dtype=np.uint32)
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_table_load_pairs(self._h, _ptr(keys), _ptr(vals), len(keys), int(can)))

    def dump(self, min_count=1, max_count=0xFFFFFFFF, sorted_=True, canonical_repr=None):
        """jellyfish dump -L min -U max -> (packed k-mers, counts)."""
        pk, pc, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        can = self.is_ds if canonical_repr is None else canonical_repr
        check(_lib.lib().tg_table_export(self._h, min_count, max_count, int(sorted_), int(can), C.byref(pk),
                                         C.byref(pc), C.byref(n)))
        try:
            keys = np.ctypeslib.as_array(C.cast(pk, C.POINTER(C.c_uint64)), shape=(max(n.value, 1),))[:n.value].copy()
            cnts = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_uint32)), shape=(max(n.value, 1),))[:n.value].copy()
        finally:
            _lib.lib().tg_free(pk)
            _lib.lib().tg_free(pc)
        return keys, cnts

    def count_sum(self):
        """sum of the counts of every k-mer in the table"""
        v = C.c_uint64(0)
        check(_lib.lib().tg_table_count_sum(self._h, C.byref(v)))
        return int(v.value)

    def histo(self):
        bins = np.zeros(_lib.TG_HISTO_BINS, dtype=np.uint64)
        check(_lib.lib().tg_histo(self._h, _ptr(bins)))
        return bins

    def coverage_stats(self, recs, offs, capture_coverage_info=False, canonical=None):
        """median/mean/stdev per record (fastaToKmerCoverageStats.cpp:300-402)."""
        recs = _as_u8(recs)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        med = np.zeros(n, dtype=np.uint32)
        mean = np.zeros(n, dtype=np.float32)
        sd = np.zeros(n, dtype=np.float32)
        per = np.zeros(int(offs[-1]) if n else 0, dtype=np.uint32) if capture_coverage_info else None
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_cov_stats(self._h, _ptr(recs), _ptr(offs), n, int(can), _ptr(med), _ptr(mean), _ptr(sd),
                                      _ptr(per)))
        return (med, mean, sd, per) if capture_coverage_info else (med, mean, sd)

    def coverage_stats_dev(self, d_recs, d_offs, nreads, d_median, d_mean, d_stdev, canonical=None):
        can = self.is_ds if canonical is None else canonical
        check(_lib.lib().tg_cov_stats_dev(self._h, d_recs, d_offs, nreads, int(can), d_median, d_mean, d_stdev))


class BundleKmerTable(_Table):
    """forward k-mer -> Inchworm-bundle index (NonRedKmerTable as ReadsToTranscripts uses it)."""
    KIND = _lib.TG_TABLE_LABEL

    def __init__(self, ctx, k=25, expected_keys=1 << 20, min_kmer_entropy=1.5):
        super().__init__(ctx, k, expected_keys)
        self.entropy_ok = np.zeros(26 * 26 * 26, dtype=np.uint8)
        _lib.lib().tg_entropy_table(k, C.c_float(min_kmer_entropy), _ptr(self.entropy_ok))

    def label_bundles(self, recs, offs, first_index=0):
        recs = _as_u8(recs)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        check(_lib.lib().tg_label_bundles(self._h, _ptr(recs), _ptr(offs), len(offs) - 1, first_index))

    def label_bundles_dev(self, d_recs, nbytes, d_offs, nbundles, first_index=0):
        check(_lib.lib().tg_label_bundles_dev(self._h, d_recs, nbytes, d_offs, nbundles, first_index))

    def assign_reads_dev(self, d_recs, d_offs, nreads, d_entropy_ok, d_best, d_pct, strand=False):
        check(_lib.lib().tg_assign_reads_dev(self._h, d_recs, d_offs, nreads, int(strand), d_entropy_ok, d_best, d_pct))

    def assign_reads(self, recs, offs, strand=False):
        """-> (best bundle index or -1, pct_read_mapped, max run) per record (ReadsToTranscripts.cc:216-274)."""
        recs = _as_u8(recs)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        best = np.zeros(n, dtype=np.int32)
        pct = np.zeros(n, dtype=np.int32)
        score = np.zeros(n, dtype=np.int32)
        check(_lib.lib().tg_assign_reads(self._h, _ptr(recs), _ptr(offs), n, int(strand), _ptr(self.entropy_ok),
                                         _ptr(best), _ptr(pct), _ptr(score)))
        return best, pct, score


class WeldmerTable:
    """NonRedKmerTable of GraphFromFasta's weldmer candidates (Chrysalis/analysis/GraphFromFasta.cc:1412-1424): a fixed set of
    kk-mers (33..48 bases), the forward windows of the reads that equal one are counted."""

    def __init__(self, ctx, weldmers, kk=48):
        blob = b"".join(w if isinstance(w, bytes) else w.encode() for w in weldmers)
        if len(blob) != len(weldmers) * kk:
            raise ValueError("every weldmer must have exactly kk characters")
        self.ctx, self.kk, self.n = ctx, kk, len(weldmers)
        self._blob = np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(0, np.uint8)
        h = C.c_void_p()
        check(_lib.lib().tg_weld_create(ctx._h, kk, _ptr(self._blob) if self.n else None, self.n, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().tg_weld_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def add_records(self, recs):
        a = _as_u8(recs)
        check(_lib.lib().tg_weld_count_reads(self._h, _ptr(a), a.nbytes))

    def add_records_dev(self, d_recs, nbytes):
        check(_lib.lib().tg_weld_count_reads_dev(self._h, d_recs, nbytes))

    def counts(self):
        out = np.zeros(self.n, np.int32)
        check(_lib.lib().tg_weld_counts(self._h, _ptr(out) if self.n else None))
        return out
