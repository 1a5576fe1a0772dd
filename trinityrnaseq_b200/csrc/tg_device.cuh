// Device-side building blocks shared by every kernel of the k-mer hot path (sm_100a only).
//
// Representation choices:
//  * bases are coded A=0 C=1 G=2 T=3, case-insensitive; everything else is "invalid" and breaks a window
//    (jellyfish rule; Inchworm contains_non_gatc, Inchworm/src/sequenceUtil.cpp:30-50;
//    Chrysalis Regular(), Chrysalis/analysis/NonRedKmerTable.cc:3-8).
//  * a read tile is held BIT-SLICED: for every chunk of 32 bases three 32-bit planes (low code bit, high
//    code bit, invalid flag), produced by three warp ballots.  A k-mer at offset o of a chunk is then two
//    funnel shifts, and its reverse complement is two bit reversals (brev of the complemented plane).
//  * the table key is the pair of k-bit planes: key = TAG | plane1 << 32 | plane0 (k <= 31).  It is a
//    bijection of the k-mer, which is all a hash table needs; the lexicographic 2-bit packing
//    (first base most significant, A<C<G<T) that the C-ABI speaks is produced/consumed by
//    planes_to_packed()/packed_to_planes() at the boundary (export, load_pairs).
//  * slot = 16 B {u64 key, u32 val, u32 aux}: key and value share one 32-B DRAM sector.
//  * MINIMIZER PLACEMENT (tg_minimizer.cuh): a key's HOME slot is slot j of the 128-B bucket chosen by the hash of its
//    minimizer, j = position of the minimizer in the key.  Consecutive windows of a read share their minimizer, so the
//    lookups of a read touch ~L/5 buckets (as runs of neighbouring slots) instead of L random DRAM granules.  A key whose
//    home slot is taken by another key raises the slot's DISPLACED flag (bit 31 of aux) and is placed by its OWN hash
//    inside the same partition: open addressing from the first slot of a 64-B group, walked linearly.  A lookup reads the
//    home slot; it goes on to the key-hashed walk only when the slot holds another key AND its flag is set.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "tg_minimizer.cuh"

namespace tg {

struct __align__(16) Slot {
    unsigned long long key;   // 0 = empty, else KEY_TAG | planes
    unsigned int val;         // count (count tables); label tables: bundle index + 1 of the k-mer that IS the key
    unsigned int aux;         // bit 31: DISPLACED flag (a key whose home is this slot lives elsewhere); low 31 bits, label
};                            // tables only: bundle index + 1 of the k-mer whose REVERSE COMPLEMENT is the key

constexpr unsigned long long KEY_TAG = 1ull << 63;
constexpr unsigned BUCKET_SLOTS = HOME_SLOTS; // 128 B home bucket; every partition is a whole number of buckets
constexpr unsigned WALK_SLOTS = 4;            // the key-hashed walk starts at a 64-B group boundary (two 256-bit loads)
constexpr unsigned long long WALK_LIMIT = 1ull << 16;   // slots an insert walks before it declares its partition full
constexpr unsigned AUX_DISPLACED = 0x80000000u;
constexpr unsigned AUX_LABEL_MASK = 0x7FFFFFFFu;
// Count tables have no use for the low 31 bits of aux, so they make the DISPLACED note selective: a key that loses its home
// slot also sets bit aux_filter_bit(key) there, and a lookup goes on to the key-hashed walk only when ITS bit is set.  The
// error variants of an expressed k-mer share its minimizer and so its home slot; most of them occur once and are not in a
// `dump -L 2` table at all -- without the filter every one of their lookups would walk for nothing.
__device__ __forceinline__ unsigned aux_filter_bit(unsigned long long key) {
    unsigned v = ((unsigned)key * 0x9E3779B1u + (unsigned)(key >> 32) * 0x85EBCA77u) >> 27;      // 0..31
    return 1u << (v == 31u ? 0u : v);
}

// Table geometry.  The table is an array of `nparts` PARTITIONS of `subcap` slots each; a key lives in the partition
// chosen by the top bits of its home hash and never leaves it (both its home bucket and its key-hashed walk are inside).
// Partitions are what make the table shardable and cache-blockable without changing a single lookup:
//   * one GPU: nparts is chosen so that a partition (subcap * 16 B) fits comfortably in L2; the partitioned
//     count path replays a k-mer log partition by partition, so its CAS/RED traffic stays in L2;
//   * N GPUs: rank r holds partitions [part0, part0 + nlocal) of the same global geometry -- owner(key) is just
//     part(key) / nlocal -- and an all-gather of the shards IS the full table (part0 = 0, nlocal = nparts).
struct Geo {
    unsigned long long subcap;   // slots per partition (a multiple of BUCKET_SLOTS)
    unsigned int nparts;         // partitions in the global table
    unsigned int part0;          // first partition held by this view
    unsigned int nlocal;         // partitions held by this view (slots[] has nlocal * subcap entries)
    int k;                       // k-mer length of the keys (the home of a key depends on it)
    unsigned int filter;         // 1 = count table: the low 31 bits of aux are a filter of the slot's displaced keys
};

struct TableView {
    Slot* slots;
    Geo g;
    unsigned long long* n_claimed;    // device counter of distinct keys
    int* error;                       // device error flag (probe overflow / key outside the local partitions)
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    // murmur3 fmix64: full avalanche (the key-hashed walk of displaced keys)
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
__device__ __forceinline__ unsigned key_p0(unsigned long long key) { return (unsigned)key; }
__device__ __forceinline__ unsigned key_p1(unsigned long long key) { return (unsigned)(key >> 32) & 0x7FFFFFFFu; }
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

// ---- addressing ----------------------------------------------------------------------------------------
// home hj (tg_minimizer.cuh: pack_home) inside this view: base = first slot of the partition, home = the slot.
// false = the partition is not held
__device__ __forceinline__ bool home_of(const Geo& g, unsigned hj, unsigned long long& base, unsigned long long& home) {
    const unsigned part = home_part(hj, g.nparts) - g.part0;
    base = (unsigned long long)part * g.subcap;
    home = base + (unsigned long long)home_bucket(hj, (unsigned)(g.subcap / BUCKET_SLOTS)) * BUCKET_SLOTS + home_slot(hj);
    return part < g.nlocal;
}
// packed home of a key from its planes alone (slow path)
__device__ __forceinline__ unsigned key_home_packed(unsigned long long key, int k);
// first slot (inside the partition) of the key-hashed walk of a displaced key
__device__ __forceinline__ unsigned long long walk_start(const Geo& g, unsigned long long key) {
    return __umul64hi(mix64(key), g.subcap / WALK_SLOTS) * WALK_SLOTS;
}
__device__ __forceinline__ unsigned long long walk_next(const Geo& g, unsigned long long off) {
    return off + 1 == g.subcap ? 0ull : off + 1;
}

// two adjacent slots (32 B, one sector) in one 256-bit load: {key, val | aux << 32} twice
__device__ __forceinline__ void ld_slot_pair(const Slot* p, unsigned long long& k0, unsigned long long& w0,
                                             unsigned long long& k1, unsigned long long& w1) {
    asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k0), "=l"(w0), "=l"(k1), "=l"(w1) : "l"(p));
}

// ---- table primitives --------------------------------------------------------------------------------
// Keys never change once written and slots never return to empty, so a stale (L1/L2) read of a key can
// only be "empty" where the truth is "claimed"; the CAS that follows re-validates.  ld.cg keeps random
// sectors out of L1.
//
// The slot of `key` (claimed if new; claimed++ then), or nullptr on a full partition / a foreign key (error raised).
// Split in two so that a caller can have the home-slot loads of several keys in flight before it settles any of them:
// `cur` is the whole home slot read with ld.cg ({key lo, key hi, val, aux}).
__device__ __forceinline__ uint4 ld_slot(const Slot* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ Slot* table_upsert_finish(const TableView& t, unsigned long long key, unsigned long long base,
                                                     unsigned long long home, uint4 cur4, unsigned& claimed) {
    Slot* hs = &t.slots[home];
    unsigned long long cur = ((unsigned long long)cur4.y << 32) | cur4.x;
    if (cur == key) return hs;
    if (cur == 0ull) {
        const unsigned long long old = atomicCAS(&hs->key, 0ull, key);
        if (old == 0ull) { claimed++; return hs; }
        if (old == key) return hs;
        cur4.w = 0u;                      // the slot was claimed under our eyes: its note cannot be trusted
    }
    // the home slot belongs to another key: leave a note there (once) and go by the key's own hash
    const unsigned note = AUX_DISPLACED | (t.g.filter ? aux_filter_bit(key) : 0u);
    if ((cur4.w & note) != note) atomicOr(&hs->aux, note);
    unsigned long long off = walk_start(t.g, key);
    // A walk is a handful of slots at the loads the host keeps (<= 0.7).  WALK_LIMIT slots without a free one means the
    // partition is full: raise the overflow flag and give up -- and once it is up every other insert gives up at once, so an
    // undersized table fails in milliseconds instead of walking millions of slots per key.
    const unsigned long long limit = t.g.subcap < WALK_LIMIT ? t.g.subcap : WALK_LIMIT;
    for (unsigned long long probes = 0; probes <= limit; probes++) {
        Slot* sl = &t.slots[base + off];
        cur = __ldcg(&sl->key);
        if (cur == key) return sl;
        if (cur == 0ull) {
            const unsigned long long old = atomicCAS(&sl->key, 0ull, key);
            if (old == 0ull) { claimed++; return sl; }
            if (old == key) return sl;
        }
        if ((probes & 63ull) == 63ull && *reinterpret_cast<volatile int*>(t.error) != 0) return nullptr;
        off = walk_next(t.g, off);
    }
    atomicExch(t.error, 1);
    return nullptr;
}
__device__ __forceinline__ Slot* table_upsert(const TableView& t, unsigned long long key, unsigned hj, unsigned& claimed) {
    unsigned long long base, home;
    if (!home_of(t.g, hj, base, home)) { atomicExch(t.error, 2); return nullptr; }
    return table_upsert_finish(t, key, base, home, ld_slot(&t.slots[home]), claimed);
}

// val += cnt (count tables)
__device__ __forceinline__ void table_add(const TableView& t, unsigned long long key, unsigned hj, unsigned cnt,
                                          unsigned& claimed) {
    Slot* s = table_upsert(t, key, hj, claimed);
    if (s) atomicAdd(&s->val, cnt);
}
__device__ __forceinline__ unsigned key_home_packed(unsigned long long key, int k) {
    unsigned h, j;
    key_home(key_p0(key), key_p1(key), k, h, j);
    return pack_home(h, j);
}
// ... for a key whose home is not known yet (slow path: 8 m-mer hashes)
__device__ __forceinline__ void table_add_key(const TableView& t, unsigned long long key, unsigned cnt, unsigned& claimed) {
    table_add(t, key, key_home_packed(key, t.g.k), cnt, claimed);
}

// Label tables (ReadsToTranscripts) are keyed by the canonical form min(k-mer, reverse complement) and keep TWO
// labels per slot: val for the bundle k-mer that equals the key, aux (low 31 bits) for the bundle k-mer whose reverse
// complement equals the key.  The reference looks every read window up twice in a table of forward strings -- once as it
// is, once reverse-complemented (ReadsToTranscripts.cc:236-250); here both answers sit in the one 16-B slot, so a
// double-stranded read costs one probe per window instead of two.  `rc` says which orientation the caller's k-mer has
// relative to the key.  Labels keep the MAXIMUM (the reference's single-thread last-writer-wins order, R4).
// Returns true when this call gave the (key, orientation) its first label, i.e. a new distinct forward k-mer.
__device__ __forceinline__ bool slot_label_max(Slot* s, bool rc, unsigned lab) {
    if (!rc) return atomicMax(&s->val, lab) == 0u;
    // aux shares its word with the DISPLACED flag: a flagged word is larger than any label, so a max against a bare
    // label is a no-op there -- repeat it among the flagged values.  Whatever the interleaving with the flag's atomicOr,
    // the word ends as flag | largest label, and exactly one caller sees "no label before mine".
    unsigned old = atomicMax(&s->aux, lab);
    if (old & AUX_DISPLACED) old = atomicMax(&s->aux, lab | AUX_DISPLACED);
    return (old & AUX_LABEL_MASK) == 0u;
}
__device__ __forceinline__ bool table_label_max(const TableView& t, unsigned long long key, unsigned hj, bool rc, unsigned lab) {
    unsigned dummy = 0;
    Slot* s = table_upsert(t, key, hj, dummy);
    if (!s || lab == 0u) return false;
    return slot_label_max(s, rc, lab);
}

// Key-hashed walk of a read-only lookup (the home slot held another key and its DISPLACED flag was set): one 64-B group
// per round, the four slots examined in fill order; a free slot ends the walk.  Returns {val, aux} or 0 (aux is only
// meaningful in label tables: in count tables it is the note of the slot the key was found in).
__device__ __forceinline__ uint2 table_walk_find(const Slot* __restrict__ slots, const Geo& g, unsigned long long base,
                                                  unsigned long long key) {
    unsigned long long off = walk_start(g, key);
    const unsigned long long limit = g.subcap < WALK_LIMIT ? g.subcap : WALK_LIMIT;     // an insert never went further
    for (unsigned long long probes = 0; probes <= limit; probes += WALK_SLOTS) {
        const Slot* b = &slots[base + off];
        unsigned long long k0, w0, k1, w1, k2, w2, k3, w3;
        ld_slot_pair(b, k0, w0, k1, w1);
        if (k0 == key) return make_uint2((unsigned)w0, (unsigned)(w0 >> 32) & AUX_LABEL_MASK);
        if (k1 == key) return make_uint2((unsigned)w1, (unsigned)(w1 >> 32) & AUX_LABEL_MASK);
        if (k0 == 0ull || k1 == 0ull) break;
        ld_slot_pair(b + 2, k2, w2, k3, w3);
        if (k2 == key) return make_uint2((unsigned)w2, (unsigned)(w2 >> 32) & AUX_LABEL_MASK);
        if (k3 == key) return make_uint2((unsigned)w3, (unsigned)(w3 >> 32) & AUX_LABEL_MASK);
        if (k2 == 0ull || k3 == 0ull) break;
        off = (off + WALK_SLOTS == g.subcap) ? 0ull : off + WALK_SLOTS;
    }
    return make_uint2(0u, 0u);
}

// Home-slot half of a lookup.  found: {val, aux} are the answer.  Otherwise `walk` says whether the key-hashed walk is
// needed (slot taken by another key and flagged) or the key is simply absent.
struct HomeProbe { uint2 v; bool found, walk; unsigned long long base; };
__device__ __forceinline__ HomeProbe table_home_find(const Slot* __restrict__ slots, const Geo& g, unsigned long long key,
                                                      unsigned hj) {
    HomeProbe r;
    unsigned long long home;
    r.v = make_uint2(0u, 0u); r.found = false; r.walk = false;
    if (!home_of(g, hj, r.base, home)) return r;                 // a shard that does not hold the partition: absent
    const uint4 s = __ldcg(reinterpret_cast<const uint4*>(&slots[home]));
    const unsigned long long sk = ((unsigned long long)s.y << 32) | s.x;
    if (sk == key) { r.v = make_uint2(s.z, g.filter ? 0u : s.w & AUX_LABEL_MASK); r.found = true; }
    else {
        const unsigned note = AUX_DISPLACED | (g.filter ? aux_filter_bit(key) : 0u);
        r.walk = sk != 0ull && (s.w & note) == note;
    }
    return r;
}
// complete single-thread lookup from a key alone (slow path: long reads, tests)
__device__ __forceinline__ uint2 table_find_key(const Slot* __restrict__ slots, const Geo& g, unsigned long long key) {
    const HomeProbe r = table_home_find(slots, g, key, key_home_packed(key, g.k));
    if (r.found || !r.walk) return r.v;
    return table_walk_find(slots, g, r.base, key);
}

// ---- k-mer log (partitioned count path) ------------------------------------------------------------------
// Phase 1 appends the k-mers of the reads to the bin of their table partition instead of touching the table; phase 2
// replays the log bin by bin, so the CAS/RED traffic of one bin stays inside an L2-resident group of partitions.
//
// An entry is a SUPER-K-MER: a run of up to 8 consecutive windows of a read that share one minimizer occurrence -- they
// all live in the same bucket of the same partition (tg_minimizer.cuh), so the run travels as ONE 16-byte entry: the
// k + n - 1 <= 32 bases as two plane words, the minimizer hash, and n / the minimizer's offset in the first window.  A
// 100-bp read is ~20 entries instead of 76 keys.  Windows whose minimizer is not unique inside the window (the same
// smallest hash twice: tandem repeats, hairpins) travel as single explicit keys with their packed home.  Homopolymer
// windows never enter the log: they are tallied per launch in hpoly[] (keys in [0..3], occurrence counts in [4..7], indexed
// by base code).
//
// Multi-GPU: the log IS the exchange.  Bin b belongs to rank b >> lp_shift (lp = bins per rank, a power of two), and
// phase 1 stores every entry straight into the OWNER's receive log through peer memory (NVLink P2P stores, the
// pointers come from CUDA IPC): owner[r] is rank r's receive log, laid out [nranks][lp][cap], and this rank writes
// segment `src` of it.  The cursors stay on the writing GPU (one atomicAdd per non-empty bin per tile, all local);
// only the entries cross NVLink, as runs of consecutive 16-byte stores, while the kernel is still rolling the next tile.
// One GPU is the same layout with one owner: owner[0] = the local log, lp_shift = 31, src = 0.
constexpr int LOG_MAX_RANKS = 8;
struct __align__(16) LogEntry {
    unsigned int b0, b1;      // run: planes of its k + n - 1 bases (base i at bit i, masked); explicit: planes of the KEY
    unsigned int h;           // run: hash of the shared minimizer; explicit: packed home hj of the key
    unsigned int meta;        // 0 = no entry; LE_VALID | n | q0 << 4 | canonical << 7 | explicit << 8
};
constexpr unsigned LE_VALID = 0x80000000u, LE_CANONICAL = 1u << 7, LE_EXPLICIT = 1u << 8;
__device__ __forceinline__ unsigned le_n(unsigned meta) { return meta & 15u; }
__device__ __forceinline__ unsigned le_q0(unsigned meta) { return (meta >> 4) & 7u; }
// most windows one run may hold: its bases must fit the 32-bit plane words
__host__ __device__ __forceinline__ int le_max_run(int k) { return 33 - k < HOME_SLOTS ? 33 - k : HOME_SLOTS; }

// window t of a run -> its table key and packed home
__device__ __forceinline__ void le_window(const LogEntry& e, unsigned t, int k, unsigned mk, unsigned long long& key,
                                          unsigned& hj) {
    const unsigned f0 = (e.b0 >> t) & mk, f1 = (e.b1 >> t) & mk;
    unsigned long long kf = make_key(f0, f1);
    bool is_rc = false;
    if (e.meta & LE_CANONICAL) {
        const unsigned long long kr = make_key(rc_plane(f0, k), rc_plane(f1, k));
        is_rc = kr < kf;
        if (is_rc) kf = kr;
    }
    const unsigned q = le_q0(e.meta) - t;                    // the minimizer's offset inside this window
    key = kf;
    hj = pack_home(e.h, is_rc ? (unsigned)(HOME_SLOTS - 1) - q : q);
}
// every k-mer of an entry added `cnt` times (the slow sides of the log: a full bin, a crowded fold table)
__device__ __forceinline__ void le_apply(const TableView& t, const LogEntry& e, unsigned cnt, unsigned& claimed) {
    const int k = t.g.k;
    if (e.meta & LE_EXPLICIT) { table_add(t, make_key(e.b0, e.b1), e.h, cnt, claimed); return; }
    const unsigned mk = kmask(k), n = le_n(e.meta);
    for (unsigned w = 0; w < n; w++) {
        unsigned long long key; unsigned hj;
        le_window(e, w, k, mk, key, hj);
        table_add(t, key, hj, cnt, claimed);
    }
}

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
