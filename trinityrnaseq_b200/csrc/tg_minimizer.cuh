// Minimizer placement: where a k-mer lives in the table is decided by its MINIMIZER, so that the consecutive windows of a
// read -- which share their minimizer for several steps -- land in ONE 128-byte bucket instead of one random DRAM
// granule each.  Everything in this file is a pure function of k-mer bits, written __host__ __device__ so that the CPU
// test-suite (tests/cpp/test_minimizer.cpp) can check the fast paths (minimizers of all windows of a read from a
// sliding minimum) against the slow one (minimizer of a single key), bit for bit, without a GPU.
//
//   * m-mer length m = k - 7: a k-mer contains w = 8 m-mers, and 8 is also the number of 16-byte slots in a 128-byte
//     bucket.  The position j (0..7) of the minimizer inside the k-mer is the slot: k-mers that share one minimizer
//     OCCURRENCE (a "super-k-mer": up to 8 consecutive windows, k + 7 = 32 bases for k = 25) differ in j, so they sit side
//     by side in the bucket without colliding, and the windows of a read touch a bucket's slots as one contiguous run.
//   * the minimizer is the m-mer with the smallest hash (ordering value = hash with its low 5 bits replaced by the
//     position, so ties go to the LEFTMOST m-mer of the key); the m-mer hash is strand-symmetric (hash of the smaller of
//     the m-mer and its reverse complement), so a k-mer and its reverse complement have the same minimizer, mirrored:
//     position q in one strand is position 7 - q in the other.  A window that is the reverse complement of its canonical
//     key therefore uses the RIGHTMOST smallest m-mer of the window, which is the leftmost one of the key.
//   * home(key) = (h, j): h = full 32-bit hash of the chosen m-mer -> partition (top bits) and bucket inside the partition
//     (remixed); j -> slot.  A key whose home slot is taken by another key sets the slot's DISPLACED flag and is placed by
//     its own hash inside the same partition (tg_device.cuh); lookups go there only when the flag is set.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TG_HD __host__ __device__ __forceinline__
#else
#define TG_HD inline
#endif

namespace tg {

constexpr int HOME_SLOTS = 8;            // slots per home bucket (128 B) = m-mers per k-mer
constexpr int MIN_FAST_K = HOME_SLOTS;   // the sliding-minimum fast paths assume exactly 8 m-mers per k-mer
constexpr unsigned ORD_POS_BITS = 5;     // low bits of the ordering value carry the position (strips of up to 32 m-mers)
constexpr unsigned ORD_POS_MASK = (1u << ORD_POS_BITS) - 1u;

TG_HD int mm_len(int k) { return k >= HOME_SLOTS ? k - (HOME_SLOTS - 1) : 1; }
TG_HD int mm_win(int k) { return k - mm_len(k) + 1; }      // m-mers per k-mer: 8 for k >= 8, k below

TG_HD unsigned bits_mask(int n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); }

TG_HD unsigned brev32(unsigned x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
// reverse complement in plane form (base i at bit i; A=00 C=01 G=10 T=11): complement flips both planes, reversal is a
// bit reversal
TG_HD unsigned rc_plane_n(unsigned p, int n) { return brev32(~p) >> (32 - n); }

// bits [o, o+32) of the 64-bit value hi:lo
TG_HD unsigned funnel_r(unsigned lo, unsigned hi, unsigned o) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, o);
#else
    o &= 31u;
    return o ? (lo >> o) | (hi << (32u - o)) : lo;
#endif
}

// strand-symmetric hash of an m-mer given as two m-bit planes: both strands are hashed (two multiply-adds each) and the
// smaller hash goes through the finalizer -- no 64-bit compare-and-select to find the canonical m-mer first
TG_HD unsigned mmer_hash(unsigned f0, unsigned f1, int m) {
    const unsigned r0 = rc_plane_n(f0, m), r1 = rc_plane_n(f1, m);
    const unsigned a = f0 * 0x9E3779B1u + f1 * 0x85EBCA77u, b = r0 * 0x9E3779B1u + r1 * 0x85EBCA77u;
    unsigned x = a < b ? a : b;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    return x;
}

// ordering values: the minimum over a range picks the smallest hash, ties to the leftmost / rightmost position
TG_HD unsigned ord_left(unsigned x, unsigned pos) { return (x & ~ORD_POS_MASK) | pos; }
TG_HD unsigned ord_right(unsigned x, unsigned pos) { return (x & ~ORD_POS_MASK) | (ORD_POS_MASK - pos); }

// home of a table key given as k-bit planes (SLOW path: 8 m-mer hashes; loaders, rehash, long reads)
TG_HD void key_home(unsigned p0, unsigned p1, int k, unsigned& h, unsigned& j) {
    const int m = mm_len(k), w = mm_win(k);
    const unsigned mm = bits_mask(m);
    unsigned best = 0xFFFFFFFFu, bh = 0;
    for (int q = 0; q < w; q++) {
        const unsigned x = mmer_hash((p0 >> q) & mm, (p1 >> q) & mm, m);
        const unsigned v = ord_left(x, (unsigned)q);
        if (v < best || q == 0) { best = v; bh = x; }
    }
    h = bh;
    j = best & ORD_POS_MASK;
}

// A home travels as ONE word: hj = (hash & ~7) | slot.  Partition and bucket are taken from a bijective REMIX of the hash with
// its low three bits cleared -- never from the hash itself: a minimizer is the SMALLEST of eight hashes, so its value is far
// from uniform (an eighth of all minimizers fall into the lowest 1/64 of the range), while its remix is uniform over the
// distinct minimizers.  Every holder of an hj (log entries, queued walks) addresses exactly like the code that computed it.
TG_HD unsigned pack_home(unsigned h, unsigned j) { return (h & ~7u) | j; }
TG_HD unsigned home_slot(unsigned hj) { return hj & 7u; }
TG_HD unsigned home_mix(unsigned hj) {
    unsigned x = (hj & ~7u) * 0x9E3779B1u;
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    return x;
}
TG_HD unsigned home_part(unsigned hj, unsigned nparts) { return (unsigned)(((unsigned long long)home_mix(hj) * nparts) >> 32); }
// (inside a partition the top bits of the remix are fixed: the bucket takes a second remix)
TG_HD unsigned home_bucket(unsigned hj, unsigned nbuckets) {
    const unsigned x = home_mix(hj) * 0xC2B2AE35u;          // the high bits of the product depend on every bit of the remix
    return (unsigned)(((unsigned long long)x * nbuckets) >> 32);
}

// ---- fast path: minimizers of the windows of a strip --------------------------------------------------------------
// hx[s], s = 0 .. PER + 6, are the m-mer hashes of a strip of PER consecutive windows (window i covers hx[i .. i + 7]).
// For every window: vl = min ord_left, vr = min ord_right over its 8 m-mers, positions relative to the strip start.
// Prefix / suffix minima around the split 7 | 8: window 0 = S[0]; window i >= 1 = min(S[i], P[i + 7]).
template <int PER>
TG_HD void strip_minimizers(const unsigned (&hx)[PER + HOME_SLOTS - 1], unsigned (&vl)[PER], unsigned (&vr)[PER]) {
    static_assert(PER >= 1 && PER <= HOME_SLOTS, "a strip holds at most 8 windows");
    unsigned sl[HOME_SLOTS], sr[HOME_SLOTS];
    sl[7] = ord_left(hx[7], 7u);
    sr[7] = ord_right(hx[7], 7u);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int s = 6; s >= 0; s--) {
        const unsigned a = ord_left(hx[s], (unsigned)s), b = ord_right(hx[s], (unsigned)s);
        sl[s] = a < sl[s + 1] ? a : sl[s + 1];
        sr[s] = b < sr[s + 1] ? b : sr[s + 1];
    }
    vl[0] = sl[0];
    vr[0] = sr[0];
    unsigned pl = 0xFFFFFFFFu, pr = 0xFFFFFFFFu;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = 1; i < PER; i++) {
        const unsigned a = ord_left(hx[i + 7], (unsigned)(i + 7)), b = ord_right(hx[i + 7], (unsigned)(i + 7));
        pl = a < pl ? a : pl;
        pr = b < pr ? b : pr;
        vl[i] = sl[i] < pl ? sl[i] : pl;
        vr[i] = sr[i] < pr ? sr[i] : pr;
    }
}
// one window on its own: hx[0..7] are the hashes of its eight m-mers
TG_HD void window_minimizers(const unsigned (&hx)[HOME_SLOTS], unsigned& vl, unsigned& vr) {
    unsigned a = ord_left(hx[0], 0u), b = ord_right(hx[0], 0u);
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int q = 1; q < HOME_SLOTS; q++) {
        const unsigned l = ord_left(hx[q], (unsigned)q), r = ord_right(hx[q], (unsigned)q);
        a = l < a ? l : a;
        b = r < b ? r : b;
    }
    vl = a; vr = b;
}
// position (relative to the strip start) of the m-mer that is the KEY's minimizer, and its slot j in key orientation;
// i = window index inside the strip, is_rc = the key is the reverse complement of the window
TG_HD unsigned strip_pick(unsigned vl, unsigned vr, int i, bool is_rc, unsigned& j) {
    const unsigned s = is_rc ? ORD_POS_MASK - (vr & ORD_POS_MASK) : (vl & ORD_POS_MASK);
    const unsigned q = s - (unsigned)i;                       // 0..7 inside the window
    j = is_rc ? (unsigned)(HOME_SLOTS - 1) - q : q;
    return s;
}

}  // namespace tg
