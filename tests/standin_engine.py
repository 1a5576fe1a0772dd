"""Stand-in engine for the CPU (gloo) tests of trinityrnaseq_b200.sharded -- TEST INFRASTRUCTURE ONLY.

`ShardedKmerCounter` drives an engine through a small interface (create_shard / new_log / partition / replay /
shard_slots / full_table / ...).  The product engine is CUDA (`sharded.DeviceEngine`).  This stand-in computes the
k-mers with the CPU oracle and keeps the shard in a dict, so that the host-side logic -- geometry, bin ownership,
the equal-split all-to-all, the [src, lp, cap] receive layout, the all-gather of slot arrays, the reductions --
runs on CPU tensors over gloo with world_size > 1.  It is never imported by the package.
"""
import numpy as np
import torch

from oracle import oracle_py as orc

MASK64 = (1 << 64) - 1


def _mix(x):
    x &= MASK64
    x ^= x >> 33; x = (x * 0xff51afd7ed558ccd) & MASK64
    x ^= x >> 33; x = (x * 0xc4ceb9fe1a85ec53) & MASK64
    x ^= x >> 33
    return x


def key_bin(key, nbins):
    return ((_mix(int(key)) & 0xFFFFFFFF) * nbins) >> 32


class StandinTable:
    def __init__(self, k, canonical, subcap, nparts, part0, nlocal):
        self.k, self.canonical = k, canonical
        self.subcap, self.nparts, self.part0, self.nlocal = subcap, nparts, part0, nlocal
        self.counts = {}

    def add(self, key, n):
        b = key_bin(key, self.nparts)
        if not (self.part0 <= b < self.part0 + self.nlocal):
            raise RuntimeError("k-mer routed to a shard that does not own its partition")
        self.counts[key] = self.counts.get(key, 0) + n

    # slot array: per partition the (key, count) pairs packed from its first slot, rest empty (key 0 = empty, so
    # keys are stored +1)
    def slots(self):
        a = np.zeros((self.nlocal, self.subcap, 2), dtype=np.int64)
        fill = [0] * self.nlocal
        for key in sorted(self.counts):
            p = key_bin(key, self.nparts) - self.part0
            a[p, fill[p]] = (key + 1, self.counts[key])
            fill[p] += 1
        return a

    def load_slots(self, a):
        a = a.reshape(self.nlocal, self.subcap, 2)
        self.counts = {}
        for p in range(self.nlocal):
            for key1, c in a[p]:
                if key1:
                    assert key_bin(int(key1) - 1, self.nparts) - self.part0 == p
                    self.counts[int(key1) - 1] = int(c)

    # what the tests compare
    def dump(self, min_count=1):
        ks = sorted(k for k, c in self.counts.items() if c >= min_count)
        return np.array(ks, dtype=np.uint64), np.array([self.counts[k] for k in ks], dtype=np.uint32)

    def size(self):
        return len(self.counts)

    def set_distinct(self, n):
        assert n == len(self.counts)

    def clear(self):
        self.counts = {}

    def coverage_stats(self, recs, offs):
        kc = orc.KmerCounter(self.k, self.canonical)
        for key, c in self.counts.items():
            kmer = "".join("ACGT"[(key >> (2 * (self.k - 1 - i))) & 3] for i in range(self.k))
            kc.add_kmer(kmer, c)
        return kc.coverage_stats(recs, offs)


class StandinPeers:
    def __init__(self, maps, rkeys, world):
        self.maps, self.rkeys, self.world, self.ok = maps, rkeys, world, True


class StandinEngine:
    """peer_dir: a directory all rank processes share; given one, the engine offers the "peer" exchange with the
    receive logs as memory-mapped files (what CUDA IPC + NVLink stores are on the device)."""

    def __init__(self, k, canonical, peer_dir=None):
        self.k, self.canonical = k, canonical
        self.table = None
        self.peer_dir = peer_dir
        self._gen = 0
        if peer_dir is None:              # no shared directory: ShardedKmerCounter must pick the collective exchange
            self.open_peer_logs = None

    def sync(self):
        pass

    def open_peer_logs(self, dist, group, nbins, cap):
        import os
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self._gen += 1
        path = lambda r: os.path.join(self.peer_dir, f"log{self._gen}_{r}.bin")
        mine = np.memmap(path(rank), dtype=np.int64, mode="w+", shape=(world, nbins // world, cap))
        mine.flush()
        dist.barrier(group=group)
        maps = [mine if r == rank else np.memmap(path(r), dtype=np.int64, mode="r+", shape=(world, nbins // world, cap))
                for r in range(world)]
        return StandinPeers(maps, torch.from_numpy(mine), world)

    def unmap_peer_logs(self, peers):
        for m in peers.maps:
            m.flush()
        peers.maps = []

    def free_peer_log(self, peers):
        peers.rkeys = None

    def new_fine_log(self, nfine, cap):
        return torch.zeros((nfine, cap), dtype=torch.int64), torch.zeros((nfine,), dtype=torch.int32)

    def refine(self, rkeys, rcur, nsrc, ncoarse, fkeys, fcur, fine0, nfine_global):
        fcur.zero_()
        nfine, fcap = fkeys.shape
        f = nfine // ncoarse
        rk = rkeys.reshape(nsrc, ncoarse, -1)
        rc = rcur.reshape(nsrc, ncoarse)
        overflow = False
        for cb in range(ncoarse):
            for s_ in range(nsrc):
                for e in rk[s_, cb, :int(rc[s_, cb])].tolist():
                    b = key_bin(e - 1, nfine_global) - fine0
                    assert cb * f <= b < (cb + 1) * f, "a coarse bin holds only its own partitions"
                    pos = int(fcur[b])
                    if pos >= fcap:                 # like the device: the entry is dropped, the caller repeats with more room
                        overflow = True
                        continue
                    fkeys[b, pos] = e
                    fcur[b] += 1
        return overflow

    def new_cursors(self, nbins):
        return (torch.zeros((nbins,), dtype=torch.int32), torch.zeros((nbins,), dtype=torch.int32),
                torch.zeros((8,), dtype=torch.int64))

    def partition_peers(self, recs, nbytes, peers, cursor, hpoly, nbins, cap, rank):
        lp = nbins // peers.world
        scratch = torch.zeros((nbins, cap), dtype=torch.int64)
        overflow = self.partition(recs, nbytes, scratch, cursor, hpoly)
        for b in range(nbins):                      # entry (bin b, pos) -> owner b // lp, segment [rank], bin b % lp
            n = int(cursor[b])
            peers.maps[b // lp][rank, b % lp, :n] = scratch[b, :n].numpy()
        for m in peers.maps:
            m.flush()
        return overflow

    def create_shard(self, subcap, nparts, part0, nlocal):
        self.table = StandinTable(self.k, self.canonical, subcap, nparts, part0, nlocal)
        return self.table

    def new_log(self, nbins, cap):
        return (torch.zeros((nbins, cap), dtype=torch.int64), torch.zeros((nbins,), dtype=torch.int32),
                torch.zeros((8,), dtype=torch.int64))

    def reset_log(self, cursor, hpoly):
        cursor.zero_()
        hpoly.zero_()

    def partition(self, recs, nbytes, keys, cursor, hpoly):
        nbins, cap = keys.shape
        ok, oc = orc.jf_count(np.asarray(recs[:nbytes]), self.k, self.canonical, 1)
        homo = {0: 0, (1 << (2 * self.k)) - 1: 3}      # A^k and T^k as packed k-mers -> side-channel slot
        overflow = False
        for key, c in zip(ok.tolist(), oc.tolist()):
            if key in homo:                         # homopolymers travel as (key, count), like on the device
                hpoly[homo[key]] = -(key + 1)       # negative, like a device key with its tag bit
                hpoly[4 + homo[key]] += c
                continue
            b = key_bin(key, nbins)
            for _ in range(c):                      # one log entry per occurrence, like the device log
                pos = int(cursor[b])
                if pos >= cap:                      # like the device: dropped, reported, the batch is repeated
                    overflow = True
                    break
                keys[b, pos] = key + 1
                cursor[b] += 1
        return overflow

    def replay(self, keys, cursor, hpoly, nsrc):
        lp = self.table.nlocal
        for i in range(4):
            n = int(hpoly[4 + i])
            if n:
                key = -int(hpoly[i]) - 1
                b = key_bin(key, self.table.nparts)
                if self.table.part0 <= b < self.table.part0 + self.table.nlocal:
                    self.table.add(key, n)
                hpoly[4 + i] = 0
        keys = keys.reshape(nsrc, lp, -1)
        cursor = cursor.reshape(nsrc, lp)
        for s in range(nsrc):
            for p in range(lp):
                for e in keys[s, p, :int(cursor[s, p])].tolist():
                    self.table.add(e - 1, 1)

    # -- routed lookups (what DeviceEngine does with k_log_tiles<QUERY>, k_query_answer, k_query_scatter, k_cov_stats) ---------
    def new_query_log(self, nbins, cap):
        return (torch.zeros((nbins, cap), dtype=torch.int64), torch.zeros((nbins,), dtype=torch.int32),
                torch.zeros((nbins, cap), dtype=torch.int32))

    def query_partition(self, recs, nbytes, keys, cursor, posidx):
        """the (canonical, packed, +1) key of every window of k bases -> the bin of its partition; posidx = where it starts"""
        nbins, cap = keys.shape
        buf = bytes(np.asarray(recs[:nbytes]).tobytes()).upper()
        code = {65: 0, 67: 1, 71: 2, 84: 3}
        k, overflow = self.k, False
        for p in range(len(buf) - k + 1):
            w = buf[p:p + k]
            if any(ch not in code for ch in w):
                continue
            f = 0
            for ch in w:
                f = (f << 2) | code[ch]
            r = 0
            for ch in reversed(w):
                r = (r << 2) | (3 - code[ch])
            key = min(f, r) if self.canonical else f
            b = key_bin(key, nbins)
            pos = int(cursor[b])
            if pos >= cap:
                overflow = True
                continue
            keys[b, pos] = key + 1
            posidx[b, pos] = p
            cursor[b] += 1
        self._qkeys = keys
        return overflow

    def query_answer(self, rkeys, rcur, resp):
        nsrc, lp, cap = rkeys.shape
        rc = rcur.reshape(nsrc, lp)
        for s_ in range(nsrc):
            for b in range(lp):
                for i in range(int(rc[s_, b])):
                    key = int(rkeys[s_, b, i]) - 1
                    assert self.table.part0 <= key_bin(key, self.table.nparts) < self.table.part0 + self.table.nlocal, \
                        "a query reached a shard that does not own its partition"
                    resp[s_, b, i] = self.table.counts.get(key, 0)

    def query_scatter_stats(self, back, posidx, cursor, recs, nbytes, offs, nreads, min_count, median, mean, stdev):
        """answers -> counts at the windows' positions -> the oracle's per-read statistics over those counts"""
        counts = np.zeros(int(nbytes) + 64, dtype=np.int64)
        nbins = back.shape[0]
        for b in range(nbins):
            n = int(cursor[b])
            counts[posidx[b, :n].numpy()] = back[b, :n].numpy()
        # the statistics of a table that holds exactly the answered k-mers at or above the floor
        kc = orc.KmerCounter(self.k, self.canonical)
        seen = {}
        for b in range(nbins):
            for i in range(int(cursor[b])):
                seen[int(self._qkeys[b, i]) - 1] = int(back[b, i])
        for key, c in seen.items():
            if c >= max(1, int(min_count)):
                kc.add_kmer("".join("ACGT"[(key >> (2 * (self.k - 1 - j))) & 3] for j in range(self.k)), c)
        m, a, sd = kc.coverage_stats(np.asarray(recs[:nbytes]), np.asarray(offs))
        median[:nreads] = m; mean[:nreads] = a; stdev[:nreads] = sd
        self.last_counts = counts

    def slots_of(self, table):
        return torch.from_numpy(table.slots().reshape(-1))

    def count_min(self, min_count):
        return sum(1 for c in self.table.counts.values() if c >= min_count)

    def new_shard_like(self, subcap):
        t = self.table
        return StandinTable(self.k, self.canonical, subcap, t.nparts, t.part0, t.nlocal)

    def compact_into(self, min_count, dst):
        dst.counts = {k: c for k, c in self.table.counts.items() if c >= min_count}

    def full_table(self, subcap, nparts):
        full = StandinTable(self.k, self.canonical, subcap, nparts, 0, nparts)
        buf = torch.zeros(nparts * subcap * 2, dtype=torch.int64)
        full._buf = buf
        orig = full.set_distinct

        def set_distinct(n, full=full, buf=buf, orig=orig):
            full.load_slots(buf.numpy())
            orig(n)
        full.set_distinct = set_distinct
        return full, buf

    def local_distinct(self):
        return self.table.size()

    def local_histo(self):
        return orc.jf_histo(self.table.dump()[1])

    def local_dump(self, min_count=1):
        return self.table.dump(min_count)

    def scalar_tensor(self, values, dtype):
        return torch.tensor(values, dtype=dtype)
