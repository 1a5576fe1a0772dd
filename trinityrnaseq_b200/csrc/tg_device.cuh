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
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace tg {

struct __align__(16) Slot {
    unsigned long long key;   // 0 = empty, else KEY_TAG | planes
    unsigned int val;         // count (count tables) or bundle index + 1 (label tables)
    unsigned int aux;         // unused (keeps the slot 16-B aligned inside one 32-B sector)
};

constexpr unsigned long long KEY_TAG = 1ull << 63;

struct TableView {
    Slot* slots;
    unsigned long long cap;           // number of slots (any size: index = mulhi64(hash, cap))
    unsigned long long* n_claimed;    // device counter of distinct keys
    int* error;                       // device error flag (probe overflow)
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

// home slot = high 64 bits of hash * capacity (any capacity, no modulo); probing is linear with wrap-around
__device__ __forceinline__ unsigned long long home_slot(unsigned long long key, unsigned long long cap) {
    return __umul64hi(mix64(key), cap);
}
__device__ __forceinline__ unsigned long long next_slot(unsigned long long idx, unsigned long long cap) {
    return (idx + 1 == cap) ? 0ull : idx + 1;
}
// owner rank of a key when the table is sharded by hash across GPUs (low hash bits; the slot uses the high ones)
__device__ __forceinline__ unsigned owner_rank(unsigned long long key, unsigned nranks) {
    return (unsigned)((mix64(key) & 0xFFFFFFFFull) * (unsigned long long)nranks >> 32);
}

// ---- table primitives --------------------------------------------------------------------------------
// Keys never change once written and slots never return to empty, so a stale (L1/L2) read of a key can
// only be "empty" where the truth is "claimed"; the CAS that follows re-validates.  ld.cg keeps random
// sectors out of L1.
__device__ __forceinline__ Slot* table_upsert_slot(const TableView& t, unsigned long long key,
                                                   unsigned long long idx, unsigned long long cur,
                                                   unsigned& claimed) {
    unsigned long long probes = 0;
    while (true) {
        if (cur == key) return &t.slots[idx];
        if (cur == 0ull) {
            unsigned long long old = atomicCAS(&t.slots[idx].key, 0ull, key);
            if (old == 0ull) { claimed++; return &t.slots[idx]; }
            if (old == key) return &t.slots[idx];
        }
        if (++probes > t.cap) { atomicExch(t.error, 1); return nullptr; }
        idx = next_slot(idx, t.cap);
        cur = __ldcg(&t.slots[idx].key);
    }
}

__device__ __forceinline__ void table_add(const TableView& t, unsigned long long key, unsigned cnt,
                                          unsigned& claimed) {
    unsigned long long idx = home_slot(key, t.cap);
    unsigned long long cur = __ldcg(&t.slots[idx].key);
    Slot* s = table_upsert_slot(t, key, idx, cur, claimed);
    if (s) atomicAdd(&s->val, cnt);
}

// read-only probe: returns val, or 0 when the key is absent.  One 16-B load fetches key and value.
__device__ __forceinline__ unsigned table_lookup(const Slot* __restrict__ slots, unsigned long long cap,
                                                 unsigned long long key) {
    unsigned long long idx = home_slot(key, cap);
    while (true) {
        const uint4 s = __ldcg(reinterpret_cast<const uint4*>(&slots[idx]));
        unsigned long long k = ((unsigned long long)s.y << 32) | s.x;
        if (k == key) return s.z;
        if (k == 0ull) return 0u;
        idx = next_slot(idx, cap);
    }
}

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
