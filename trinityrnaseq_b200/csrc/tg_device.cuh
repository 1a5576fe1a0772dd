// Device-side building blocks shared by every kernel of the k-mer hot path (sm_100a only).
//
// Representation choices:
//  * bases are coded A=0 C=1 G=2 T=3, case-insensitive; everything else is "invalid" and breaks a window
//    (jellyfish rule; Inchworm contains_non_gatc, Inchworm/src/sequenceUtil.cpp:30-50;
//    Chrysalis Regular(), Chrysalis/analysis/NonRedKmerTable.cc:3-8).
//  * a read tile is held BIT-SLICED: for every chunk of 32 bases three 32-bit planes (low code bit, high
//    code bit, invalid flag), produced by three warp ballots.  A k-mer at offset o of a chunk is then two
//    funnel shifts, and its reverse complement is two bit reversals (brev of the complemented plane).
//  * the table key is the pair of k-bit planes: key = TAG | plane1 << 32 | plane0 (k <= 31; k = 32 fills all 64 bits,
//    has no TAG, and its one k-mer that reads as "empty" -- poly-A, key 0 -- lives in a slot of its own behind the
//    partitions: see zero_key_slot).  It is a
//    bijection of the k-mer, which is all a hash table needs; the lexicographic 2-bit packing
//    (first base most significant, A<C<G<T) that the C-ABI speaks is produced/consumed by
//    planes_to_packed()/packed_to_planes() at the boundary (export, load_pairs).
//  * slot = 16 B {u64 key, u32 val, u32 aux}: key and value share one 32-B DRAM sector.
//  * BUCKETS: a key's probe sequence starts at the first slot of a 64-B bucket (4 slots, one DRAM access granule:
//    ncu shows ~80 B of DRAM traffic per random 16-B load, i.e. the neighbours come along anyway) and then walks
//    linearly.  Inserts claim the first free slot of that walk, so a bucket fills front to back and a lookup settles
//    a key with ONE round of two 256-bit loads (LDG.E.ENL2.256) unless the whole bucket is taken -- the dependent
//    probe rounds that a warp-convergent lookup pays as max-over-lanes shrink from 3-4 to ~1.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace tg {

struct __align__(16) Slot {
    unsigned long long key;   // 0 = empty, else KEY_TAG | planes
    unsigned int val;         // count (count tables); label tables: bundle index + 1 of the k-mer that IS the key
    unsigned int aux;         // label tables: bundle index + 1 of the k-mer whose REVERSE COMPLEMENT is the key
};                            // (count tables leave aux 0; it keeps the slot 16-B aligned inside one 32-B sector)

constexpr unsigned long long KEY_TAG = 1ull << 63;
constexpr unsigned BUCKET_SLOTS = 4;          // 64 B; every partition is a whole number of buckets

// Table geometry.  The table is an array of `nparts` PARTITIONS of `subcap` slots each; a key lives in
// partition part(key) (low hash word) and probes linearly, wrapping INSIDE its partition (slot from the high hash
// word).  Partitions are what make the table shardable and cache-blockable without changing a single lookup:
//   * one GPU: nparts is chosen so that a partition (subcap * 16 B) fits comfortably in L2; the partitioned
//     count path replays a k-mer log partition by partition, so its CAS/RED traffic stays in L2;
//   * N GPUs: rank r holds partitions [part0, part0 + nlocal) of the same global geometry -- owner(key) is just
//     part(key) / nlocal -- and an all-gather of the shards IS the full table (part0 = 0, nlocal = nparts).
struct Geo {
    unsigned long long subcap;   // slots per partition
    unsigned int nparts;         // partitions in the global table
    unsigned int part0;          // first partition held by this view
    unsigned int nlocal;         // partitions held by this view (slots[] has nlocal * subcap entries)
    int k;                       // k-mer length of the keys
    unsigned int floor;          // count tables, read paths: a count below this reads as absent (a `dump -L floor` VIEW of the table)
};

struct TableView {
    Slot* slots;
    Geo g;
    unsigned long long* n_claimed;    // device counter of distinct keys
    int* error;                       // device error flag (probe overflow / key outside the local partitions)
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    // murmur3 fmix64: full avalanche, so low bits index the table and high bits pick the owner GPU
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__host__ __device__ __forceinline__ unsigned base_code(unsigned c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
__host__ __device__ __forceinline__ bool base_valid(unsigned c) {
    unsigned u = c & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

__device__ __forceinline__ unsigned kmask(int k) { return k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u); }

__device__ __forceinline__ unsigned long long make_key(unsigned p0, unsigned p1) {
    return KEY_TAG | ((unsigned long long)p1 << 32) | p0;
}
// k = 32 (Inchworm/src/KmerCounter.cpp:15-17 allows it): the two planes take all 64 bits, so the key carries no tag.  Every
// k-mer but one still differs from the empty marker 0; poly-A (both planes 0) does not, and is kept OUTSIDE the probed
// range: slot [nlocal * subcap], the first of the spare bucket every table is allocated with.  The kernels that accept
// k = 32 build their keys with make_key_k and reach that slot through the `key == 0` branches of table_update /
// table_label_max / table_lookup2; they are the flat-tile, pair-loading, CTA-per-read and scan kernels.  The warp-per-read
// and log kernels (the measured path, k <= 31) keep the constant tag and never see a zero key.
__device__ __forceinline__ unsigned long long make_key_k(unsigned p0, unsigned p1, int k) {
    return (k < 32 ? KEY_TAG : 0ull) | ((unsigned long long)p1 << 32) | p0;
}
// reverse complement in plane form: complement flips both code bits, reversal is a bit reversal
__device__ __forceinline__ unsigned rc_plane(unsigned p, int k) { return __brev(~p) >> (32 - k); }

// planes (base i at bit i) -> lexicographic 2-bit packing (base 0 most significant)
__host__ __device__ __forceinline__ unsigned long long planes_to_packed(unsigned p0, unsigned p1, int k) {
    unsigned long long v = 0;
    for (int i = 0; i < k; i++) {
        unsigned c = ((p0 >> i) & 1u) | (((p1 >> i) & 1u) << 1);
        v = (v << 2) | c;
    }
    return v;
}
__host__ __device__ __forceinline__ void packed_to_planes(unsigned long long v, int k, unsigned& p0, unsigned& p1) {
    p0 = 0; p1 = 0;
    for (int i = 0; i < k; i++) {
        unsigned c = (unsigned)(v >> (2 * (k - 1 - i))) & 3u;
        p0 |= (c & 1u) << i;
        p1 |= (c >> 1) << i;
    }
}
__host__ __device__ __forceinline__ unsigned long long packed_revcomp(unsigned long long v, int k) {
    unsigned long long r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (3ull - (v & 3ull)); v >>= 2; }
    return r;
}

// Key hash: two 32-bit words from ~16 32-bit instructions (a 64-bit murmur finalizer costs twice that in IMADs, and every
// window of every read pays it).  x picks the PARTITION (and so the owner GPU and the log bin), y the bucket inside it;
// both go through the full murmur3 fmix32 (consecutive windows of a read are shifted copies of each other: a weaker mix
// sends neighbours to correlated bins, which the shared-memory counters of phase 1 feel), from different combinations of
// the key's plane words, so together they carry ~64 bits.
__device__ __forceinline__ unsigned fmix32(unsigned x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    return x ^ (x >> 16);
}
struct KeyHash { unsigned x, y; };
__device__ __forceinline__ KeyHash key_hash(unsigned long long key) {
    const unsigned a = (unsigned)key * 0x9E3779B1u, b = (unsigned)(key >> 32) * 0x85EBCA77u;
    KeyHash h;
    h.x = fmix32(a + b);
    h.y = fmix32(a ^ __funnelshift_l(b, b, 15));
    return h;
}
// partition of a hash, any partition count.  Nested: with nfine = f * ncoarse, hash_part(h, nfine) / f == hash_part(h, ncoarse)
__device__ __forceinline__ unsigned hash_part(KeyHash h, unsigned nparts) { return __umulhi(h.x, nparts); }
struct Probe {
    unsigned long long base;   // first slot of the key's partition inside this view
    unsigned long long off;    // current slot inside the partition
};
// returns false when the key's partition is not held by this view
__device__ __forceinline__ bool probe_home(const Geo& g, unsigned long long key, Probe& p) {
    const KeyHash h = key_hash(key);
    const unsigned part = hash_part(h, g.nparts) - g.part0;
    p.base = (unsigned long long)part * g.subcap;
    p.off = (unsigned long long)__umulhi(h.y, (unsigned)(g.subcap / BUCKET_SLOTS)) * BUCKET_SLOTS;      // first slot of the home bucket
    return part < g.nlocal;
}
__device__ __forceinline__ void probe_next(const Geo& g, Probe& p) { p.off = (p.off + 1 == g.subcap) ? 0ull : p.off + 1; }
// home of the one key that equals the empty marker (poly-A at k = 32): behind the last partition of an unsharded table
template <typename S>
__device__ __forceinline__ S* zero_key_slot(S* slots, const Geo& g) { return slots + (unsigned long long)g.nlocal * g.subcap; }

// ---- table primitives --------------------------------------------------------------------------------
// Keys never change once written and slots never return to empty, so a stale (L1/L2) read of a key can
// only be "empty" where the truth is "claimed"; the CAS that follows re-validates.  ld.cg keeps random
// sectors out of L1.
__device__ __forceinline__ Slot* table_upsert_slot(const TableView& t, unsigned long long key, Probe p,
                                                   unsigned long long cur, unsigned& claimed) {
    unsigned long long probes = 0;
    while (true) {
        Slot* sl = &t.slots[p.base + p.off];
        if (cur == key) return sl;
        if (cur == 0ull) {
            unsigned long long old = atomicCAS(&sl->key, 0ull, key);
            if (old == 0ull) { claimed++; return sl; }
            if (old == key) return sl;
        }
        if (++probes > t.g.subcap) { atomicExch(t.error, 1); return nullptr; }
        probe_next(t.g, p);
        cur = __ldcg(&t.slots[p.base + p.off].key);
    }
}

// val += cnt (count tables) or val = max(val, cnt) (label tables)
template <bool IS_MAX>
__device__ __forceinline__ void table_update(const TableView& t, unsigned long long key, unsigned v, unsigned& claimed) {
    if (key == 0ull) {              // poly-A at k = 32: no claim, the slot is its own (a first non-zero value = a new key)
        Slot* z = zero_key_slot(t.slots, t.g);
        const unsigned old = IS_MAX ? atomicMax(&z->val, v) : atomicAdd(&z->val, v);
        if (old == 0u && v != 0u) claimed++;
        return;
    }
    Probe p;
    if (!probe_home(t.g, key, p)) { atomicExch(t.error, 2); return; }
    const unsigned long long cur = __ldcg(&t.slots[p.base + p.off].key);
    Slot* s = table_upsert_slot(t, key, p, cur, claimed);
    if (s) { if (IS_MAX) atomicMax(&s->val, v); else atomicAdd(&s->val, v); }
}

// Label tables (ReadsToTranscripts) are keyed by the canonical form min(k-mer, reverse complement) and keep TWO
// labels per slot: val for the bundle k-mer that equals the key, aux for the bundle k-mer whose reverse complement
// equals the key.  The reference looks every read window up twice in a table of forward strings -- once as it is,
// once reverse-complemented (ReadsToTranscripts.cc:236-250); here both answers sit in the one 16-B slot, so a
// double-stranded read costs one probe per window instead of two (and the second one was nearly always a miss
// that had to walk to an empty slot).  `rc` says which orientation the caller's k-mer has relative to the key.
// Returns true when this call gave the (key, orientation) its first label, i.e. a new distinct forward k-mer.
__device__ __forceinline__ bool table_label_max(const TableView& t, unsigned long long key, bool rc, unsigned lab) {
    if (key == 0ull) {              // poly-A at k = 32 (rc: the bundle holds poly-T)
        Slot* z = zero_key_slot(t.slots, t.g);
        return lab != 0u && atomicMax(rc ? &z->aux : &z->val, lab) == 0u;
    }
    Probe p;
    if (!probe_home(t.g, key, p)) { atomicExch(t.error, 2); return false; }
    const unsigned long long cur = __ldcg(&t.slots[p.base + p.off].key);
    unsigned dummy = 0;
    Slot* s = table_upsert_slot(t, key, p, cur, dummy);
    if (!s || lab == 0u) return false;
    return atomicMax(rc ? &s->aux : &s->val, lab) == 0u;
}

// two adjacent slots (32 B, one sector) in one 256-bit load: {key, val | aux << 32} twice
__device__ __forceinline__ void ld_slot_pair(const Slot* p, unsigned long long& k0, unsigned long long& w0,
                                             unsigned long long& k1, unsigned long long& w1) {
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(w0), "=l"(k1), "=l"(w1) : "l"(p));
}

// The same through the read-only path (L1-cached: ld.global.nc).  Only for tables that no kernel in flight writes to -- the
// per-read kernels; in locus order neighbouring reads ask for the same buckets, and L1 answers many of them.
__device__ __forceinline__ void ld_slot_pair_ro(const Slot* p, unsigned long long& k0, unsigned long long& w0,
                                                unsigned long long& k1, unsigned long long& w1) {
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(w0), "=l"(k1), "=l"(w1) : "l"(p));
}

// Read-only probe: returns {val, aux}, or 0 when the key is absent or !valid.  One round = one 64-B bucket.
// CONVERGENT: every lane of the warp must call it (lanes without a key pass valid = false).  Lanes leave the probe
// loops at different times; the __syncwarp() between the hot-table phase and the big-table phase brings them back
// together, so the DRAM-bound loads of a warp are issued as one request and not once per straggler group.
__device__ __forceinline__ uint2 table_lookup2(const Slot* __restrict__ slots, const Geo& g, unsigned long long key,
                                                bool valid) {
    const KeyHash h = key_hash(key);
    uint2 v = make_uint2(0u, 0u);   // {val, aux}
    bool open = valid;
    const unsigned part = hash_part(h, g.nparts) - g.part0;
    if (part >= g.nlocal) open = false;
    if (valid && key == 0ull) {     // poly-A at k = 32: its own slot behind the partitions, nothing to compare
        const uint4 z = __ldcg(reinterpret_cast<const uint4*>(zero_key_slot(slots, g)));
        v = make_uint2(z.z, z.w);
        open = false;
    }
    const unsigned long long base = (unsigned long long)part * g.subcap;
    unsigned long long off = (unsigned long long)__umulhi(h.y, (unsigned)(g.subcap / BUCKET_SLOTS)) * BUCKET_SLOTS;
    // one bucket per round, the four slots examined in fill order
    for (unsigned long long probes = 0; open && probes <= g.subcap; probes += BUCKET_SLOTS) {   // bounded: a full partition cannot hang
        const Slot* b = &slots[base + off];
        unsigned long long k0, w0, k1, w1, k2, w2, k3, w3;
        ld_slot_pair(b, k0, w0, k1, w1);
        ld_slot_pair(b + 2, k2, w2, k3, w3);
        // Short-circuit on purpose: ptxas sinks the second load behind the outcome of the first pair, so a lookup settled by
        // slots 0-1 (97 % in a count table at load 0.34) requests ONE sector.  Measured the other way (both halves always
        // requested, branch-free match): k_cov_stats 35 -> 58 ms -- the second request is not free even though DRAM
        // delivers the whole 64 B.  In a `dump -L 2` table a third of the lookups (absent singletons, slots 2-3) pay a
        // second dependent round: 43 ms instead of 35 ms.  (Also measured and dropped: the lookups of a whole read -- three
        // groups of 32 windows -- in flight together; it needs 64 registers, and the lost occupancy costs more than the
        // saved round trips: 35 -> 38-40 ms.  The kernel is throughput-bound: 69 % issue slots, 66 % of DRAM peak.)
        unsigned long long w = 0ull;
        bool hit = true;
        if (k0 == key) w = w0; else if (k1 == key) w = w1; else if (k2 == key) w = w2; else if (k3 == key) w = w3; else hit = false;
        if (hit) { v = make_uint2((unsigned)w, (unsigned)(w >> 32)); open = false; }
        else if (k0 == 0ull || k1 == 0ull || k2 == 0ull || k3 == 0ull) open = false;     // a free slot ends the walk
        off = (off + BUCKET_SLOTS == g.subcap) ? 0ull : off + BUCKET_SLOTS;
    }
    __syncwarp();
    return v;
}
// The same lookup in two halves, so that a caller can have the first sectors of SEVERAL keys in flight before it settles
// any of them (the per-read kernels issue the loads of three rounds of 32 windows back to back: one memory latency per
// three rounds instead of three).  lookup_issue: home bucket address + its first sector; lookup_settle: the answer.
// CONVERGENT like table_lookup2: every lane of the warp calls both (lanes without a key pass valid = false); the rare
// continuations (second sector of the bucket, further buckets) are entered by the whole warp on a vote, so the common
// path has no divergence bookkeeping at all.
struct LookupIssue { const Slot* bucket; unsigned long long k0, w0, k1, w1; };
__device__ __forceinline__ LookupIssue lookup_issue(const Slot* __restrict__ slots, const Geo& g, unsigned long long key,
                                                    bool& valid) {
    LookupIssue q;
    const KeyHash h = key_hash(key);
    const unsigned part = hash_part(h, g.nparts) - g.part0;
    if (part >= g.nlocal) valid = false;                 // a shard that does not hold the partition: absent
    const unsigned long long off = (unsigned long long)__umulhi(h.y, (unsigned)(g.subcap / BUCKET_SLOTS)) * BUCKET_SLOTS;
    q.bucket = &slots[(unsigned long long)part * g.subcap + off];
    q.k0 = q.w0 = q.k1 = q.w1 = 0ull;
    if (valid) ld_slot_pair_ro(q.bucket, q.k0, q.w0, q.k1, q.w1);
    return q;
}
__device__ __forceinline__ uint2 lookup_settle(const Slot* __restrict__ slots, const Geo& g, unsigned long long key, bool valid,
                                                const LookupIssue& q) {
    uint2 v = make_uint2(0u, 0u);   // {val, aux}
    bool open = valid;
    if (open) {
        if (q.k0 == key) { v = make_uint2((unsigned)q.w0, (unsigned)(q.w0 >> 32)); open = false; }
        else if (q.k1 == key) { v = make_uint2((unsigned)q.w1, (unsigned)(q.w1 >> 32)); open = false; }
        else if (q.k0 == 0ull || q.k1 == 0ull) open = false;                      // a free slot ends the walk
    }
    if (__any_sync(0xFFFFFFFFu, open)) {
        unsigned long long k2 = 0ull, w2 = 0ull, k3 = 0ull, w3 = 0ull;
        if (open) {
            ld_slot_pair_ro(q.bucket + 2, k2, w2, k3, w3);
            if (k2 == key) { v = make_uint2((unsigned)w2, (unsigned)(w2 >> 32)); open = false; }
            else if (k3 == key) { v = make_uint2((unsigned)w3, (unsigned)(w3 >> 32)); open = false; }
            else if (k2 == 0ull || k3 == 0ull) open = false;
        }
        // whole bucket taken by other keys (rare): walk on, bucket by bucket, wrapping inside the partition
        const KeyHash h = key_hash(key);
        const unsigned long long base = (unsigned long long)(hash_part(h, g.nparts) - g.part0) * g.subcap;
        unsigned long long off = (unsigned long long)__umulhi(h.y, (unsigned)(g.subcap / BUCKET_SLOTS)) * BUCKET_SLOTS;
        for (unsigned long long probes = BUCKET_SLOTS; __any_sync(0xFFFFFFFFu, open) && probes <= g.subcap; probes += BUCKET_SLOTS) {
            off = (off + BUCKET_SLOTS == g.subcap) ? 0ull : off + BUCKET_SLOTS;      // (bounded: a full partition cannot hang)
            if (open) {
                const Slot* b = &slots[base + off];
                unsigned long long k0, w0, k1, w1;
                ld_slot_pair_ro(b, k0, w0, k1, w1);
                ld_slot_pair_ro(b + 2, k2, w2, k3, w3);
                unsigned long long w = 0ull;
                bool hit = true;
                if (k0 == key) w = w0; else if (k1 == key) w = w1; else if (k2 == key) w = w2; else if (k3 == key) w = w3; else hit = false;
                if (hit) { v = make_uint2((unsigned)w, (unsigned)(w >> 32)); open = false; }
                else if (k0 == 0ull || k1 == 0ull || k2 == 0ull || k3 == 0ull) open = false;
            }
        }
    }
    return v;
}
__device__ __forceinline__ unsigned table_lookup(const Slot* __restrict__ slots, const Geo& g, unsigned long long key,
                                                 bool valid) {
    return table_lookup2(slots, g, key, valid).x;
}

// ---- k-mer log (partitioned count path) ------------------------------------------------------------------
// Phase 1 appends every counted k-mer occurrence to the bin of its hash partition instead of touching the table;
// phase 2 replays the log bin by bin, so the CAS/RED traffic of one bin stays inside an L2-resident group of
// partitions.  An entry is the table key (0 = no entry).  Homopolymer windows never enter the log: they are
// tallied per launch in hpoly[] (keys in [0..3], occurrence counts in [4..7], indexed by base code).
//
// Multi-GPU: the log IS the exchange.  Bin b belongs to rank b >> lp_shift (lp = bins per rank, a power of two), and
// phase 1 stores every entry straight into the OWNER's receive log through peer memory (NVLink P2P stores, the
// pointers come from CUDA IPC): owner[r] is rank r's receive log, laid out [nranks][lp][cap], and this rank writes
// segment `src` of it.  The cursors stay on the writing GPU (one atomicAdd per non-empty bin per tile, all local);
// only the 8-B keys cross NVLink, as runs of consecutive entries, while the kernel is still rolling the next tile.
// One GPU is the same layout with one owner: owner[0] = the local log, lp_shift = 31, src = 0.
constexpr int LOG_MAX_RANKS = 8;
using LogEntry = unsigned long long;            // one counted occurrence = the table key (8 B)
constexpr int MIN_FAST_K = 1;                   // every k the key layout holds goes through the warp-path kernels
__host__ __device__ __forceinline__ int le_max_run(int) { return 1; }      // k-mers per log entry
struct LogView {
    LogEntry* owner[LOG_MAX_RANKS];             // receive log of each rank: [nranks][lp][cap]
    unsigned int* cursor;       // [nbins] entries reserved so far, LOCAL; may run past cap (readers clamp, writers
                                //         past cap insert directly / raise the overflow flag)
    unsigned int nbins;         // bins == partitions of the geometry the log was laid out for
    unsigned int cap;           // entries per bin (per source segment)
    unsigned int lp_shift;      // log2(bins per rank); 31 on one GPU
    unsigned int src;           // this rank = the segment it writes in every owner's log
    int* error;                 // device flag raised (3) when a bin overflows and there is no table to fall back to
    unsigned long long* hpoly;  // [8] homopolymer side channel
    unsigned int* posidx;       // QUERY logs only: [nbins][cap] LOCAL, the record-buffer position of the window whose key went
                                // to (bin, pos) -- the return address of the routed lookup
};

// ---- TMA (1-D bulk async copy) + mbarrier wrappers -----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                              unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// x86-64 SSE default NaN (sign bit set): what 0.0f/0.0f yields in the reference build, printed "-nan"
constexpr unsigned X86_DEFAULT_NAN_BITS = 0xFFC00000u;

}  // namespace tg
