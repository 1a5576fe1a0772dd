"""Hash-sharded k-mer counting across the GPUs of one box: one process per GPU, torch.distributed for
the plumbing, libtrinity_gpu for every k-mer.

  reads  shard by input offset: rank r counts the records starting in [r*S/n, (r+1)*S/n)  (`record_range`)
  table  shards by hash: the global table is `nparts = world * lp` partitions; rank r holds partitions
         [r*lp, (r+1)*lp) (tg_table_create_sharded).  Prior art for owner = f(canonical k-mer) mod n:
         Inchworm/src/mpi_deprecated/MPIinchworm.cpp:1236-1257 -- there one blocking MPI_Send per k-mer (:519-531).

  count     phase 1 on every rank appends each k-mer occurrence to the log bin of its partition; phase 2 replays the
            received bins into the shard (tg_table_replay_log_dev).  The k-mers travel in COARSE bins -- `cbins =
            world * c` of them (c = exchange_bins(): 128 in total by default), rank d owning bins [d*c, (d+1)*c) --
            because phase 1 and the NVLink stores are only fast with long runs per bin; the owner then splits each
            coarse bin into its lp / c table partitions (tg_log_refine_dev) and, from FOLD_FROM_WORLD GPUs on, folds a
            chunk's duplicate k-mers in shared memory before touching the table.  Between phase 1 and the owner the
            bins must travel:
              exchange="peer"        (default on GPUs) phase 1 IS the exchange: the kernel stores every entry straight
                                     into segment [rank] of the OWNER's receive log through peer memory (CUDA IPC,
                                     NVLink P2P stores; tg_count_partition_peers_dev), so the transfer overlaps the
                                     rolling of the next tile and there is no send buffer; only the [c] cursor rows
                                     are exchanged afterwards (a few KB, and the "all stores have landed" point);
              exchange="collective"  phase 1 fills a local send log (tg_count_partition_dev) and ONE equal-split
                                     all-to-all moves the bins (NCCL over NVLink) -- the fallback when peer memory
                                     cannot be mapped, and what the CPU test-suite runs over gloo.
  queries   the shards are all-gathered once into a full replica per GPU (`replicate`), because the
            concatenation of the shards' slot arrays IS the full table; coverage statistics / lookups then run
            locally with no per-batch communication (SURVEY §8e "all-gather once" branch).  min_count = 2 gathers the
            device-side `jellyfish dump -L 2` of every shard instead (a quarter of the bytes, identical statistics).

The exchange logic is independent of where the k-mers are computed: `ShardedKmerCounter` drives an *engine*.  The
product engine is `DeviceEngine` (CUDA).  The CPU test-suite drives the same class over gloo with a stand-in
engine defined in tests/ -- there is no CPU engine in this package.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .api import KmerCounter

TARGET_LOAD = 0.45
MAX_BINS = 8192


def record_range(offs_or_total, rank, world):
    """Records [r0, r1) of rank `rank`: those whose first byte lies in [rank*S/world, (rank+1)*S/world), S = total
    bytes.  `offs_or_total` is the offs array of a record buffer (len n+1)."""
    offs = np.asarray(offs_or_total, dtype=np.uint64)
    total = int(offs[-1])
    lo = total * rank // world
    hi = total * (rank + 1) // world
    r0 = int(np.searchsorted(offs[:-1], lo, side="left"))
    r1 = int(np.searchsorted(offs[:-1], hi, side="left"))
    return r0, r1


def shard_geometry(world, expected_keys_per_rank, part_bytes=16 << 20):
    """-> (slots_per_partition, nparts, lp): lp partitions per rank, each about part_bytes (the L2 blocking unit)."""
    slots = max(int(expected_keys_per_rank / TARGET_LOAD) + 1, 4096)
    lp = 1
    while lp * world * 2 <= MAX_BINS and slots * 16 / lp > part_bytes:
        lp *= 2
    subcap = (slots + lp - 1) // lp
    return subcap, world * lp, lp


MAX_EXCHANGE_BINS = 128       # measured at 8 GPUs: partition+exchange 26.4 ms with 128 bins, 32.4 ms with 256
FOLD_FROM_WORLD = 4           # replay folds a chunk's duplicate k-mers in shared memory from this many GPUs on: every GPU
                              # sends its copies of the same expressed k-mers to the one owner, and same-address atomics
                              # serialise in L2 (8 GPUs: replay 47.2 -> 24.9 ms; on 1-2 GPUs the fold costs more than it saves)


def exchange_bins(world, lp, max_bins=MAX_EXCHANGE_BINS):
    """-> c, the COARSE bins per rank the k-mers are exchanged in (a power of two dividing lp, world * c <= max_bins):
    phase 1 and the NVLink stores are only efficient while a tile of reads scatters into a few hundred bins; the owner
    splits each coarse bin into its lp // c table partitions afterwards (tg_log_refine_dev)."""
    c = 1
    while c * 2 <= lp and world * c * 2 <= max_bins:
        c *= 2
    return c


ENTRIES_PER_WINDOW = 1.0      # a log entry is one k-mer occurrence


def log_capacity(nbytes, nbins, slack=1.2):
    """entries per log bin for a batch of nbytes record bytes (every byte starts at most one window = one entry); the
    additive term covers the hash fluctuation of small batches, the factor covers hot k-mers.  A bin that overflows all
    the same makes the caller double the head-room and repeat."""
    cap = int(nbytes * ENTRIES_PER_WINDOW / nbins * slack) + 1024
    return (cap + 15) // 16 * 16


class DeviceEngine:
    """libtrinity_gpu behind the engine interface; buffers are torch CUDA tensors (torch = device memory + NCCL)."""

    def __init__(self, ctx, k, canonical):
        import torch
        self.torch = torch
        self.ctx, self.k, self.canonical = ctx, k, bool(canonical)
        self.device = torch.device("cuda", ctx.device)
        self.table = None

    # -- table shard ------------------------------------------------------------------------------------
    def create_shard(self, subcap, nparts, part0, nlocal):
        self.table = KmerCounter.sharded(self.ctx, self.k, self.canonical, subcap, nparts, part0, nlocal)
        return self.table

    @property
    def ENTRY_WORDS(self):
        """int64 words of one log entry on the device (tg_log_entry_bytes: the 8-byte table key)"""
        return max(1, int(_lib.lib().tg_log_entry_bytes()) // 8)

    def new_log(self, nbins, cap):
        t = self.torch
        keys = t.empty((nbins, cap, self.ENTRY_WORDS), dtype=t.int64, device=self.device)
        cursor = t.zeros((nbins,), dtype=t.int32, device=self.device)
        hpoly = t.zeros((8,), dtype=t.int64, device=self.device)    # homopolymer side channel: 4 keys, 4 counts
        return keys, cursor, hpoly

    def reset_log(self, cursor, hpoly):
        cursor.zero_()
        hpoly.zero_()
        self.torch.cuda.current_stream(self.device).synchronize()

    def partition(self, d_recs, nbytes, keys, cursor, hpoly):
        """phase 1: record buffer in HBM -> log bins"""
        nbins, cap = keys.shape[0], keys.shape[1]
        check(_lib.lib().tg_count_partition_dev(self.ctx._h, d_recs, nbytes, self.k, int(self.canonical), nbins, cap,
                                                C.c_void_p(keys.data_ptr()), C.c_void_p(cursor.data_ptr()),
                                                C.c_void_p(hpoly.data_ptr())))
        return self.ctx.log_overflow_check()          # (a sync: the exchange runs on torch's stream)

    # -- routed lookups ------------------------------------------------------------------------------------
    def new_query_log(self, nbins, cap):
        t = self.torch
        return (t.empty((nbins, cap), dtype=t.int64, device=self.device), t.zeros((nbins,), dtype=t.int32, device=self.device),
                t.empty((nbins, cap), dtype=t.int32, device=self.device))

    def query_partition(self, d_recs, nbytes, keys, cursor, posidx):
        """requester: key of every valid window -> the bin of its owner partition; posidx = the windows' positions"""
        nbins, cap = keys.shape
        check(_lib.lib().tg_query_partition_dev(self.ctx._h, d_recs, nbytes, self.k, int(self.canonical), nbins, cap,
                                                C.c_void_p(keys.data_ptr()), C.c_void_p(cursor.data_ptr()),
                                                C.c_void_p(posidx.data_ptr())))
        return self.ctx.log_overflow_check()

    def query_answer(self, rkeys, rcur, resp):
        """owner: received keys [nsrc, lp, cap] -> resp [nsrc, lp, cap] (int32: the shard's count, 0 = absent)"""
        nsrc, lp, cap = rkeys.shape
        self.torch.cuda.current_stream(self.device).synchronize()
        check(_lib.lib().tg_query_answer_dev(self.table._h, C.c_void_p(rkeys.data_ptr()), C.c_void_p(rcur.data_ptr()), nsrc, lp, cap,
                                             C.c_void_p(resp.data_ptr())))
        self.ctx.sync()

    def query_scatter_stats(self, back, posidx, cursor, d_recs, nbytes, d_offs, nreads, min_count, d_median, d_mean, d_stdev):
        """requester: answers [nbins, cap] -> counts at the windows' positions -> per-read statistics"""
        t = self.torch
        nbins, cap = back.shape
        # (+ a tile: the partition kernel works on whole tiles, and windows that start past nbytes -- bytes of the caller's next
        #  slice, or padding -- are looked up too; their answers land here and nobody reads them)
        counts = t.zeros((int(nbytes) + 16384,), dtype=t.int32, device=self.device)
        t.cuda.current_stream(self.device).synchronize()
        check(_lib.lib().tg_query_scatter_dev(self.ctx._h, C.c_void_p(back.data_ptr()), C.c_void_p(posidx.data_ptr()),
                                              C.c_void_p(cursor.data_ptr()), nbins, cap, C.c_void_p(counts.data_ptr())))
        check(_lib.lib().tg_cov_stats_counts_dev(self.ctx._h, d_recs, d_offs, nreads, self.k, int(min_count),
                                                 C.c_void_p(counts.data_ptr()), d_median, d_mean, d_stdev))
        self.ctx.sync()

    def sync(self):
        self.ctx.sync()

    def open_peer_logs(self, dist, group, nbins, cap):
        """Collective.  This rank's receive log [world, lp, cap] allocated with cudaMalloc, its IPC handle exchanged,
        every other rank's log mapped into this process.  -> PeerLogs, or None (on every rank) when any rank fails."""
        L, t = _lib.lib(), self.torch
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = nbins * cap * 8 * self.ENTRY_WORDS
        mine, handle, ok = None, b"", 1
        try:
            mine = self.ctx.dev_alloc(nbytes)
            h = (C.c_uint8 * _lib.TG_IPC_HANDLE_BYTES)()
            check(L.tg_ipc_export(self.ctx._h, mine, h))
            handle = bytes(h)
        except Exception:
            ok = 0
        handles = [None] * world
        dist.all_gather_object(handles, (ok, handle), group=group)
        ptrs, opened = (C.c_void_p * world)(), []
        if all(o for o, _ in handles):
            try:
                for r, (_, hb) in enumerate(handles):
                    if r == rank:
                        ptrs[r] = mine.value
                        continue
                    q = C.c_void_p()
                    check(L.tg_ipc_open(self.ctx._h, (C.c_uint8 * _lib.TG_IPC_HANDLE_BYTES).from_buffer_copy(hb), C.byref(q)))
                    ptrs[r] = q.value
                    opened.append(q)
            except Exception:
                ok = 0
        else:
            ok = 0
        flag = self.scalar_tensor([ok], t.int64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        peers = PeerLogs(mine, ptrs, opened, None, world)
        peers.ok = int(flag.item()) == 1
        if peers.ok:
            peers.rkeys = _alias_tensor(t, mine.value, nbytes, self.device).view(t.int64).view(world, nbins // world, cap,
                                                                                                self.ENTRY_WORDS)
        return peers

    # closing is two steps with a barrier between them (the caller's): every importer unmaps before any exporter frees
    def unmap_peer_logs(self, peers):
        for q in peers.opened:
            try:
                check(_lib.lib().tg_ipc_close(self.ctx._h, q))
            except Exception:
                pass
        peers.opened = []

    def free_peer_log(self, peers):
        if peers.mine is not None:
            peers.rkeys = None
            self.ctx.dev_free(peers.mine)
            peers.mine = None

    def partition_peers(self, d_recs, nbytes, peers, cursor, hpoly, nbins, cap, rank):
        """phase 1 fused with the exchange: entries stored into the owners' receive logs over NVLink"""
        check(_lib.lib().tg_count_partition_peers_dev(self.ctx._h, d_recs, nbytes, self.k, int(self.canonical), nbins, cap,
                                                      peers.world, rank, peers.ptrs, C.c_void_p(cursor.data_ptr()),
                                                      C.c_void_p(hpoly.data_ptr())))
        return self.ctx.log_overflow_check()          # kernel end = this rank's peer stores are visible system-wide

    def new_fine_log(self, nfine, cap):
        t = self.torch
        return (t.empty((nfine, cap, self.ENTRY_WORDS), dtype=t.int64, device=self.device),
                t.zeros((nfine,), dtype=t.int32, device=self.device))

    def refine(self, rkeys, rcur, nsrc, ncoarse, fkeys, fcur, fine0, nfine_global):
        """received coarse log [nsrc, ncoarse, cap] -> fine log [nfine, fcap], one segment per local partition"""
        fcur.zero_()
        self.torch.cuda.current_stream(self.device).synchronize()
        check(_lib.lib().tg_log_refine_dev(self.ctx._h, C.c_void_p(rkeys.data_ptr()), C.c_void_p(rcur.data_ptr()), nsrc,
                                           ncoarse, rkeys.shape[-2], C.c_void_p(fkeys.data_ptr()),
                                           C.c_void_p(fcur.data_ptr()), fkeys.shape[0], fkeys.shape[1], fine0, nfine_global))
        return self.ctx.log_overflow_check()

    def new_cursors(self, nbins):
        t = self.torch
        return (t.zeros((nbins,), dtype=t.int32, device=self.device), t.zeros((nbins,), dtype=t.int32, device=self.device),
                t.zeros((8,), dtype=t.int64, device=self.device))

    def set_replay_fold(self, on):
        self.ctx.set("replay_fold", int(bool(on)))

    def replay(self, keys, cursor, hpoly, nsrc):
        """phase 2: received log [nsrc, lp, cap] (+ global homopolymer tallies) -> this rank's shard"""
        cap = keys.shape[-2]
        self.torch.cuda.current_stream(self.device).synchronize()
        check(_lib.lib().tg_table_replay_log_dev(self.table._h, C.c_void_p(keys.data_ptr()),
                                                 C.c_void_p(cursor.data_ptr()), C.c_void_p(hpoly.data_ptr()), nsrc, cap))

    def slots_of(self, table):
        """a table's slot array as a flat torch uint8 tensor aliasing the table memory"""
        p, n = C.c_void_p(), C.c_uint64()
        check(_lib.lib().tg_table_slots_dev(table._h, C.byref(p), C.byref(n)))
        return _alias_tensor(self.torch, p.value, n.value, self.device)

    def count_min(self, min_count):
        return self.table.count_min(min_count)

    def new_shard_like(self, subcap):
        _, nparts, part0, nlocal = self.table.geometry()
        return KmerCounter.sharded(self.ctx, self.k, self.canonical, subcap, nparts, part0, nlocal)

    def compact_into(self, min_count, dst):
        self.table.compact_into(min_count, dst)
        self.ctx.sync()

    def full_table(self, subcap, nparts):
        full = KmerCounter.sharded(self.ctx, self.k, self.canonical, subcap, nparts, 0, nparts)
        return full, self.slots_of(full)

    def local_distinct(self):
        return self.table.size()

    def local_histo(self):
        return self.table.histo()

    def local_dump(self, min_count=1):
        return self.table.dump(min_count=min_count)

    def scalar_tensor(self, values, dtype):
        return self.torch.tensor(values, dtype=dtype, device=self.device)


class PeerLogs:
    """receive logs of all ranks as seen from this process (engine-specific pointers) + this rank's own as a tensor"""

    def __init__(self, mine, ptrs, opened, rkeys, world):
        self.mine, self.ptrs, self.opened, self.rkeys, self.world = mine, ptrs, opened, rkeys, world
        self.ok = True


class _CudaAlias:
    """__cuda_array_interface__ view of raw device memory, so torch can wrap a libtrinity_gpu allocation"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _alias_tensor(torch, ptr, nbytes, device):
    return torch.as_tensor(_CudaAlias(ptr, nbytes), device=device)


class ShardedKmerCounter:
    """KmerCounter whose table is sharded by hash over the ranks of a torch.distributed process group."""

    def __init__(self, engine, expected_keys_per_rank, group=None, part_bytes=16 << 20, dist=None, exchange="auto",
                 max_exchange_bins=MAX_EXCHANGE_BINS, replay_fold=None):
        if dist is None:
            import torch.distributed as dist
        if exchange not in ("auto", "peer", "collective"):
            raise ValueError("exchange must be auto, peer or collective")
        self.dist, self.group, self.eng = dist, group, engine
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.subcap, self.nparts, self.lp = shard_geometry(self.world, expected_keys_per_rank, part_bytes)
        self.table = engine.create_shard(self.subcap, self.nparts, self.rank * self.lp, self.lp)
        if exchange == "auto":
            exchange = "peer" if (getattr(engine, "open_peer_logs", None) is not None and self.world <= 8) else "collective"
        self.exchange = exchange
        # k-mers travel in cbins = world * c coarse bins; rank r owns coarse bins [r*c, (r+1)*c) = partitions [r*lp, (r+1)*lp)
        self.c = exchange_bins(self.world, self.lp, max_exchange_bins)
        self.cbins = self.world * self.c
        self._fine = None         # (fine keys [lp, cap], fine cursors [lp]) when lp > c
        if replay_fold is None:
            replay_fold = self.world >= FOLD_FROM_WORLD
        if hasattr(engine, "set_replay_fold"):
            engine.set_replay_fold(replay_fold)
        self.replay_fold = bool(replay_fold)
        self.profile = None       # set to a dict to collect host-clock milliseconds per phase (syncs around each)
        self._log = None
        self._recv = None
        self._grow = 1            # per-bin head-room multiplier: doubled whenever a batch overflowed a bin (hot k-mers)
        self._fine_grow = 1
        self.overflow_retries = 0
        self._peers = None        # (PeerLogs, cap, (cursor, rcursor, hpoly))
        self._compact = None      # (compacted shard, its slots per partition)
        self._full = None         # (full replica, its slot bytes as a tensor, slots per partition)

    def clear(self):
        self.table.clear()

    def owner_of_bin(self, b):
        """rank that owns table partition b"""
        return b // self.lp

    def _timed(self, name, fn):
        if self.profile is None:
            return fn()
        import time
        self._sync_all()
        t0 = time.perf_counter()
        out = fn()
        self._sync_all()
        self.profile[name] = self.profile.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    def _sync_all(self):
        if hasattr(self.eng, "sync"):
            self.eng.sync()
        if hasattr(self.eng, "torch"):
            self.eng.torch.cuda.current_stream(self.eng.device).synchronize()

    def _agree_capacity(self, nbytes):
        # every rank lays its log out for the largest batch (nbytes bounds the entries of a batch: every byte starts
        # at most one window).  The reduction doubles as the barrier the peer exchange needs: a rank enters it only
        # after its own previous replay has finished, so nobody is still reading a log that is about to be rewritten.
        if hasattr(self.eng, "sync"):
            self.eng.sync()
        m = self.eng.scalar_tensor([int(nbytes)], _int64(self.eng))
        self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX, group=self.group)
        self._bound = int(m.item())
        return log_capacity(self._bound, self.cbins) * self._grow

    def _peer_buffers(self, nbytes):
        cap = self._agree_capacity(nbytes)
        if self._peers is not None and self._peers[1] < cap:
            self._close_peers(self._peers[0])              # every rank is past its last replay (barrier above)
            self._peers = None
        if self._peers is None:
            peers = self.eng.open_peer_logs(self.dist, self.group, self.cbins, cap)
            if not peers.ok:
                self._close_peers(peers)
                return None
            self._peers = (peers, cap, self.eng.new_cursors(self.cbins))
        else:
            cur, _, hpoly = self._peers[2]
            self.eng.reset_log(cur, hpoly)
        return self._peers

    def _buffers(self, nbytes):
        cap = self._agree_capacity(nbytes)
        if self._log is None or self._log[0].shape[1] < cap:
            self._log = self.eng.new_log(self.cbins, cap)
            self._recv = self.eng.new_log(self.cbins, cap)       # same bytes, viewed [world, c, cap]
        else:
            self.eng.reset_log(self._log[1], self._log[2])
        return self._log, self._recv

    def _any(self, flag):
        """collective OR of a per-rank flag"""
        f = self.eng.scalar_tensor([1 if flag else 0], _int64(self.eng))
        self.dist.all_reduce(f, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(f.item()) != 0

    def add_records_dev(self, d_recs, nbytes, max_windows=None):
        """Count every k-mer of this rank's record buffer into the sharded table (collective: all ranks call it).
        max_windows: an upper bound on the k-mer windows of the buffer when the caller knows one tighter than
        nbytes (fixed-length reads: nreads * (L - k + 1)); it only sizes the exchange buffers.

        Bins are laid out for an even spread plus head-room.  A k-mer hot enough to overfill its bin on
        any rank -- an adapter dimer in a tenth of the reads, say -- does not fail the batch: the ranks agree that a bin
        overflowed, double the head-room (kept for later batches) and repeat the partition step; nothing has touched the
        table at that point."""
        bound = nbytes if max_windows is None else min(nbytes, max_windows)
        while True:
            if self.exchange == "peer":
                pb = self._timed("buffers", lambda: self._peer_buffers(bound))
                if pb is None:
                    self.exchange = "collective"          # peer memory unavailable (agreed by all ranks): fall back for good
            if self.exchange == "peer":
                peers, cap, (cur, rcur, hpoly) = pb
                # phase 1 + exchange in one kernel: this rank's entries land in segment [rank] of every owner's log
                ovf = self._timed("partition+exchange", lambda: self.eng.partition_peers(d_recs, nbytes, peers, cur, hpoly,
                                                                                        self.cbins, cap, self.rank))
            else:
                (keys, cur, hpoly), (rkeys, rcur, _) = self._timed("buffers", lambda: self._buffers(bound))
                ovf = self._timed("partition", lambda: self.eng.partition(d_recs, nbytes, keys, cur, hpoly))
            if not self._any(ovf):
                break
            if self._grow >= 64:
                raise RuntimeError("k-mer log bin still overflows at 64x head-room")
            self._grow *= 2
            self.overflow_retries += 1
        if self.exchange == "peer":
            # cursor rows [d*lp, (d+1)*lp) go to rank d (a few KB); completing it also means every rank is past its
            # kernel, i.e. all peer stores into this rank's log have landed
            self._timed("cursors", lambda: self.dist.all_to_all_single(rcur, cur, group=self.group))
            rkeys = peers.rkeys
        else:
            # bins [d*lp, (d+1)*lp) go to rank d: an equal-split all-to-all over dim 0
            self._timed("cursors", lambda: self.dist.all_to_all_single(rcur, cur, group=self.group))
            self._timed("exchange", lambda: self.dist.all_to_all_single(rkeys, keys, group=self.group))
        # homopolymer tallies: every rank learns the global counts, the owner of each key applies it
        self.dist.all_reduce(hpoly[:4], op=self.dist.ReduceOp.MIN, group=self.group)   # keys carry bit 63: negative
        self.dist.all_reduce(hpoly[4:], op=self.dist.ReduceOp.SUM, group=self.group)
        if self.lp > self.c:
            # the owner splits its coarse bins into table partitions; the source segments merge on the way
            while True:
                fcap = log_capacity(self._bound, self.lp, slack=1.3) * self._fine_grow
                if self._fine is None or self._fine[0].shape[1] < fcap:
                    self._fine = self.eng.new_fine_log(self.lp, fcap)
                fkeys, fcur = self._fine
                ovf = self._timed("refine", lambda: self.eng.refine(rkeys, rcur, self.world, self.c, fkeys, fcur,
                                                                    self.rank * self.lp, self.nparts))
                if not self._any(ovf):
                    break
                if self._fine_grow >= 64:
                    raise RuntimeError("fine k-mer log bin still overflows at 64x head-room")
                self._fine_grow *= 2
                self.overflow_retries += 1
            self._timed("replay", lambda: self.eng.replay(fkeys, fcur, hpoly, 1))
        else:
            self._timed("replay", lambda: self.eng.replay(rkeys, rcur, hpoly, self.world))

    def close(self):
        """Collective: unmap / free the peer logs (the other buffers are ordinary tensors)."""
        if self._peers is not None:
            self._agree_capacity(0)                      # barrier: nobody still writes into or replays from a log
            self._close_peers(self._peers[0])
            self._peers = None

    def _close_peers(self, peers):
        self.eng.unmap_peer_logs(peers)
        self._agree_capacity(0)                          # barrier: every importer has unmapped
        self.eng.free_peer_log(peers)

    def size(self):
        """distinct k-mers in the global table"""
        t = self.eng.scalar_tensor([self.eng.local_distinct()], _int64(self.eng))
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    def histo(self):
        """jellyfish histo of the global table = sum of the shards' histograms"""
        h = self.eng.local_histo()
        t = self.eng.scalar_tensor(np.asarray(h, dtype=np.int64).tolist(), _int64(self.eng))
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().astype(np.uint64)

    def dump_local(self, min_count=1):
        """this rank's part of `jellyfish dump` (sorted); the global dump is the merge of all ranks' parts"""
        return self.eng.local_dump(min_count)

    def coverage_stats_routed_dev(self, d_recs, nbytes, d_offs, nreads, d_median, d_mean, d_stdev, min_count=1):
        """Per-read coverage statistics of this rank's reads against the SHARDED table, without a replica (collective): the
        key of every window goes to the rank that owns its partition (8 B), the count comes back (4 B), and the statistics
        are computed from the counts.  For tables that do not fit as replicas; bit-identical to the replica path.  A bin
        that overflows (hot k-mers) makes all ranks double the head-room and repeat, like the count does."""
        grow = 1
        while True:
            m = self.eng.scalar_tensor([int(nbytes)], _int64(self.eng))
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX, group=self.group)
            cap = log_capacity(int(m.item()), self.cbins) * grow
            keys, cur, posidx = self.eng.new_query_log(self.cbins, cap)
            ovf = self._timed("q.partition", lambda: self.eng.query_partition(d_recs, nbytes, keys, cur, posidx))
            if not self._any(ovf):
                break
            if grow >= 64:
                raise RuntimeError("query log bin still overflows at 64x head-room")
            grow *= 2
        rkeys, rcur, resp = self.eng.new_query_log(self.cbins, cap)
        # bins [d*c, (d+1)*c) go to rank d; what arrives is [source rank][c][cap]
        self._timed("q.exchange", lambda: (self.dist.all_to_all_single(rcur, cur, group=self.group),
                                           self.dist.all_to_all_single(rkeys, keys, group=self.group)))
        self._timed("q.answer", lambda: self.eng.query_answer(rkeys.view(self.world, self.c, cap), rcur, resp.view(self.world, self.c, cap)))
        back = self.eng.new_query_log(self.cbins, cap)[2]
        self._timed("q.return", lambda: self.dist.all_to_all_single(back, resp, group=self.group))
        self._timed("q.stats", lambda: self.eng.query_scatter_stats(back, posidx, cur, d_recs, nbytes, d_offs, nreads, min_count,
                                                                    d_median, d_mean, d_stdev))

    def replicate(self, min_count=1, load=TARGET_LOAD):
        """All-gather the shards into a full table on every rank -> KmerCounter for local queries.  min_count > 1
        gathers only the k-mers with at least that count (`jellyfish dump -L min_count`, what the normalisation
        pipeline feeds fastaToKmerCoverageStats): coverage statistics are bit-identical for min_count <= 2 because
        they clamp counts below 1 to 1, and the replica is several times smaller.  Buffers are cached, so calling
        this once per step costs one compaction kernel and one all-gather."""
        if min_count <= 1:
            src, subcap = self.table, self.subcap
        else:
            # the compacted shard is sized once (one streaming count, agreed over ranks) and then refilled; a refill
            # that outgrows it is caught by the fill check below and sized again
            for attempt in (0, 1):
                if self._compact is None or self._compact[2] != min_count:
                    n = self.eng.scalar_tensor([self.eng.count_min(min_count)], _int64(self.eng))
                    self.dist.all_reduce(n, op=self.dist.ReduceOp.MAX, group=self.group)
                    need = max(int(int(n.item()) / load / self.lp) + 64, 64)
                    self._compact = (self.eng.new_shard_like(need), need, min_count)
                src, subcap, _ = self._compact
                self._timed("compact", lambda: self.eng.compact_into(min_count, src))
                full_flag = 0
                try:
                    if src.size() > 0.8 * subcap * self.lp:
                        full_flag = 1
                except Exception:           # overflow flag raised by the table
                    full_flag = 1
                f = self.eng.scalar_tensor([full_flag], _int64(self.eng))
                self.dist.all_reduce(f, op=self.dist.ReduceOp.MAX, group=self.group)
                if int(f.item()) == 0:
                    break
                if hasattr(src, "close"):
                    src.close()
                self._compact = None
        if self._full is None or self._full[2] != subcap:
            full, full_bytes = self.eng.full_table(subcap, self.nparts)
            self._full = (full, full_bytes, subcap)
        full, full_bytes, _ = self._full
        mine = self.eng.slots_of(src)
        self._timed("allgather", lambda: self.dist.all_gather_into_tensor(full_bytes, mine, group=self.group))
        if hasattr(self.eng, "torch"):
            self.eng.torch.cuda.current_stream(self.eng.device).synchronize()
        d = self.eng.scalar_tensor([src.size()], _int64(self.eng))
        self.dist.all_reduce(d, group=self.group)
        full.set_distinct(int(d.item()))
        return full


def _int64(eng):
    if hasattr(eng, "torch"):
        return eng.torch.int64
    import torch
    return torch.int64
