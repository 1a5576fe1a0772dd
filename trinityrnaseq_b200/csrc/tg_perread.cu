// Per-read kernels of the Trinity k-mer hot path (sm_100a):
//
//   k_cov_stats[_long]    compute_kmer_coverage / median / mean / stDev   
//   k_assign[_long]       ReadsToTranscripts per-read vote                (§8a R5, R7-R9)
//   k_read_locus          one 32-bit LOCUS signature per read (smallest strand-symmetric m-mer hash of the read)
//
// Warp path (reads up to PR_MAXWIN windows): one warp per read; in round i lane l handles window 32 i + l.
//   1. SWAR transpose of the read into bit planes (shared memory), four bases per lane and round;
//   2. per window: canonical key -> home bucket -> one warp-convergent bucket lookup (tg_device.cuh);
//   3. statistics straight from the registers that hold the lane's coverage values.
// Statistics need the sequential fp32 sum of squares (the reference's evaluation order is observable in the last bits,
// S9).  A warp keeps the coverage vectors of a BATCH of consecutive reads in its shared-memory arena and runs the
// sequential sums of the whole batch at once, one read per lane: the 76 dependent adds of a 100-bp read are issued once
// per batch instead of once per read.
//
// LOCUS ORDER.  Both kernels take an optional permutation `order`: the i-th read processed is read order[i], results go
// to the read's own index.  The callers pass the reads sorted by k_read_locus.  Reads that overlap the same stretch of a
// transcript share their smallest m-mer with high probability, so they become neighbours in that order, and an RNA-seq
// library covers every expressed position tens to thousands of times: the table buckets one read touches are the ones
// its neighbours touched microseconds earlier, and the lookups are served by L2 (and L1) instead of one random DRAM
// granule each (measured: profiles/README.md).  The order changes nothing but the timing -- every read is still looked
// up window by window in the same table.
//
// Long path (CTA per read, planes and buffers in global scratch): same per-window code, bitonic sort for the median.
#include "tg_internal.h"

namespace tg {

namespace {
constexpr unsigned FULL = 0xFFFFFFFFu;
}

// =========================================================================================================
// shared pieces
// =========================================================================================================
template <int GS> __device__ __forceinline__ void gsync() {
    if (GS == 32) __syncwarp(); else __syncthreads();
}

// planes for chunks 0..nch (chunk nch and everything past L is invalid)
template <int GS>
__device__ __forceinline__ void pack_read_planes(const uint8_t* __restrict__ seq, int L, int nch, uint32_t* P0,
                                                 uint32_t* P1, uint32_t* PB, int gtid) {
    const int lane = gtid & 31, w = gtid >> 5;
    for (int c = w; c <= nch; c += GS / 32) {
        const int pos = c * 32 + lane;
        const unsigned ch = pos < L ? seq[pos] : (unsigned)'\n';
        const unsigned code = base_code(ch);
        const unsigned b0 = __ballot_sync(FULL, code & 1u);
        const unsigned b1 = __ballot_sync(FULL, code >> 1);
        const unsigned bb = __ballot_sync(FULL, !base_valid(ch));
        if (lane == 0) { P0[c] = b0; P1[c] = b1; PB[c] = bb; }
    }
}

// ascending bitonic sort of buf[0..n2), n2 a power of two
template <int GS, typename T>
__device__ __forceinline__ void bitonic_sort(T* buf, unsigned n2, int gtid) {
    for (unsigned kk = 2; kk <= n2; kk <<= 1) {
        for (unsigned j = kk >> 1; j > 0; j >>= 1) {
            for (unsigned i = gtid; i < n2; i += GS) {
                const unsigned ixj = i ^ j;
                if (ixj > i) {
                    const T x = buf[i], y = buf[ixj];
                    const bool up = (i & kk) == 0;
                    if ((x > y) == up) { buf[i] = y; buf[ixj] = x; }
                }
            }
            gsync<GS>();
        }
    }
}

__device__ __forceinline__ unsigned next_pow2(unsigned n) {
    unsigned p = 1;
    while (p < n) p <<= 1;
    return p;
}

template <int GS>
__device__ __forceinline__ unsigned long long group_sum_u64(unsigned long long v, unsigned long long* red, int gtid) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (GS == 32) return v;
    gsync<GS>();
    if ((gtid & 31) == 0) red[gtid >> 5] = v;
    gsync<GS>();
    unsigned long long tot = 0;
    for (int w = 0; w < GS / 32; w++) tot += red[w];
    gsync<GS>();
    return tot;
}

// The warp path's transpose: FOUR bases per lane and round (128 bases per round: one round for a 100-bp read instead of
// four ballot rounds).  Lane l holds bases 4l .. 4l+3 of the round as one 32-bit word; code bits and validity are computed
// for the four bytes at once (SWAR), gathered into nibbles, and the eight lanes that share a 32-base plane word OR their
// nibbles together (xor butterfly).  Bytes are read as aligned words (two per lane, funnel-shifted by the record's
// misalignment); nothing past the record's terminator is touched beyond the padding every device record buffer carries.
__device__ __forceinline__ unsigned swar_zero_bytes(unsigned z) {          // bit 7 of every byte of z that is 0
    return ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u;
}
__device__ __forceinline__ unsigned swar_gather(unsigned x) {             // bits 0, 8, 16, 24 -> bits 0..3
    return ((x & 0x01010101u) * 0x01020408u) >> 24;
}
// OR of three words over the 8 lanes that share a plane word: an xor butterfly (a redux.sync over a partial mask is
// executed group by group -- four serialised collectives per word, measured at 10 % of the statistics kernel)
__device__ __forceinline__ void or_over_8_lanes(unsigned& a, unsigned& b, unsigned& c) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        a |= __shfl_xor_sync(FULL, a, o);
        b |= __shfl_xor_sync(FULL, b, o);
        c |= __shfl_xor_sync(FULL, c, o);
    }
}
__device__ __forceinline__ void pack_read_planes_warp(const uint8_t* __restrict__ seq, int L, int nch, uint32_t* P0,
                                                      uint32_t* P1, uint32_t* PB, int lane) {
    const unsigned long long a0 = reinterpret_cast<unsigned long long>(seq);
    const int sh = (int)(a0 & 3ull);
    const unsigned* words = reinterpret_cast<const unsigned*>(a0 - (unsigned long long)sh);
    // chunks 0 .. nch-1 hold the read.  (A window's funnel shifts may also READ word nch -- never more: the arrays are sized
    // for it -- but every bit they take from it is masked off: a window ends inside the read.)
    for (int g = 0; 4 * g < nch; g++) {
        const int first = 128 * g + 4 * lane;                               // index of this lane's first base
        // aligned words covering bytes [first - sh, first - sh + 8)
        const unsigned w0 = first - sh < L ? words[32 * g + lane] : 0x0A0A0A0Au;
        const unsigned w1 = first - sh + 4 < L ? words[32 * g + lane + 1] : 0x0A0A0A0Au;
        const unsigned w = __funnelshift_r(w0, w1, 8 * sh);
        const unsigned cc = (w >> 1) ^ (w >> 2);                            // code = bits 0-1 of every byte
        const unsigned u = w & 0xDFDFDFDFu;                                 // upper case
        const unsigned ok = swar_zero_bytes(u ^ 0x41414141u) | swar_zero_bytes(u ^ 0x43434343u) |
                            swar_zero_bytes(u ^ 0x47474747u) | swar_zero_bytes(u ^ 0x54545454u);
        const int live = L - first;                                         // bases of this lane inside the record
        const unsigned lm = live >= 4 ? 0xFu : live <= 0 ? 0u : (1u << live) - 1u;
        const unsigned n0 = swar_gather(cc), n1 = swar_gather(cc >> 1);
        const unsigned nb = ~(swar_gather(ok >> 7) & lm) & 0xFu;           // invalid: not ACGT, or past the record
        const int pos = 4 * (lane & 7);
        unsigned b0 = n0 << pos, b1 = n1 << pos, bb = nb << pos;
        or_over_8_lanes(b0, b1, bb);
        if ((lane & 7) == 0) { const int c = 4 * g + (lane >> 3); P0[c] = b0; P1[c] = b1; PB[c] = bb; }
    }
}

// planes of one warp's current read
struct WarpFront {
    uint32_t p0[PR_MAXCH], p1[PR_MAXCH], pb[PR_MAXCH];
};

__device__ __forceinline__ void front_planes(WarpFront& f, const uint8_t* __restrict__ seq, int L, int lane) {
    pack_read_planes_warp(seq, L, (L + 31) >> 5, f.p0, f.p1, f.pb, lane);
    __syncwarp();
}

// one window: planes -> canonical key (or the forward one) and orientation
struct Window { unsigned long long key; unsigned f0, f1; bool valid, is_rc, pal; };
__device__ __forceinline__ Window front_window(const WarpFront& f, int p, int nwin, int k, unsigned mk, bool canonical) {
    Window w;
    w.valid = false; w.is_rc = false; w.pal = false; w.key = 0ull; w.f0 = 0u; w.f1 = 0u;
    if (p < nwin) {
        const int c = p >> 5, o = p & 31;
        if (!(__funnelshift_r(f.pb[c], f.pb[c + 1], o) & mk)) {
            w.f0 = __funnelshift_r(f.p0[c], f.p0[c + 1], o) & mk;
            w.f1 = __funnelshift_r(f.p1[c], f.p1[c + 1], o) & mk;
            const unsigned long long kf = make_key(w.f0, w.f1), kr = make_key(rc_plane(w.f0, k), rc_plane(w.f1, k));
            w.is_rc = canonical && kr < kf;
            w.pal = kr == kf;
            w.key = w.is_rc ? kr : kf;
            w.valid = true;
        }
    }
    return w;
}

// =========================================================================================================
// coverage statistics
// =========================================================================================================
constexpr int ST_BATCH = 16;          // reads per batch at most (the arena -- coverage words per warp, a launch parameter -- may hold fewer)
constexpr int ST_RPW = 16;            // consecutive reads handled by one warp

struct StatsWarp {
    WarpFront f;
    uint32_t b_off[ST_BATCH], b_n[ST_BATCH], b_r[ST_BATCH];
    float b_avg[ST_BATCH];
    uint32_t cov[1];                  // [arena], the warp's coverage vectors (dynamic shared memory behind the header)
};
__host__ __device__ inline size_t stats_warp_bytes(int arena) { return (sizeof(StatsWarp) + (size_t)(arena - 1) * 4 + 15) / 16 * 16; }

// Median of the n values a warp holds in registers (lane l owns x[i] = value 32 i + l; elements past n hold 0xFFFFFFFF,
// which no pivot below the maximum reaches -- if every value IS 0xFFFFFFFF the answer is that value either way), WITHOUT
// sorting: a bisection on the VALUE between the warp minimum and maximum (coverage values of one read sit in a narrow
// band, so a handful of rounds), each round one compare per element and one redux.sync.  Returns median_coverage() of
// fastaToKmerCoverageStats.cpp:337-347: odd n -> the middle element, even n -> the (wrapping) u32 mean of the two middle
// elements.
template <int PER>
__device__ __forceinline__ uint32_t warp_median_regs(const unsigned (&x)[PER], int n, unsigned lo, unsigned hi) {
    const unsigned k1 = (unsigned)(n - 1) / 2u, k2 = (unsigned)n / 2u;
    // smallest value with at least k1 + 1 elements <= it = the element of rank k1
    // (measured: three pivots per round in one redux -- half the rounds -- is no faster: 25.8 vs 24.9 ms)
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        unsigned cnt = 0;
#pragma unroll
        for (int i = 0; i < PER; i++) cnt += x[i] <= mid ? 1u : 0u;          // (dead elements are 0xFFFFFFFF > mid)
        cnt = __reduce_add_sync(FULL, cnt);
        if (cnt >= k1 + 1u) hi = mid; else lo = mid + 1u;
    }
    const unsigned x1 = lo;
    if (k1 == k2) return x1;
    unsigned le = 0, nxt = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        le += x[i] <= x1 ? 1u : 0u;
        if (x[i] > x1) nxt = min(nxt, x[i]);
    }
    le = __reduce_add_sync(FULL, le);
    nxt = __reduce_min_sync(FULL, nxt);
    const unsigned x2 = le >= k2 + 1u ? x1 : nxt;          // the element of rank k2 = k1 + 1
    return (uint32_t)(x1 + x2) / 2u;
}

// lookups of one read + sum, mean, median; the coverage vector is left in sw.cov[used .. used + nwin).  The first-sector
// loads of LK rounds (LK x 32 windows) are issued back to back before any of them is examined.
template <int PER, bool PRE>
__device__ __forceinline__ void stats_read(StatsWarp& sw, const Slot* __restrict__ slots, const Geo& geo, int nwin, int k,
                                           unsigned mk, bool canonical, unsigned used, int lane, uint32_t& median, float& mean,
                                           const uint32_t* __restrict__ pre) {
    constexpr int LK = PER <= 3 ? 3 : 2;           // rounds in flight (registers: longer reads keep more values)
    unsigned x[PER];
    unsigned mn = 0xFFFFFFFFu, mx = 0u;
    unsigned long long part = 0;
#pragma unroll
    for (int i0 = 0; i0 < PER; i0 += LK) {
        unsigned long long key[LK];
        bool ok[LK];
        LookupIssue q[LK];
        if (!PRE) {
#pragma unroll
            for (int u = 0; u < LK; u++)
                if (i0 + u < PER) {
                    const Window w = front_window(sw.f, 32 * (i0 + u) + lane, nwin, k, mk, canonical);
                    key[u] = w.key; ok[u] = w.valid;
                    q[u] = lookup_issue(slots, geo, key[u], ok[u]);
                }
        }
#pragma unroll
        for (int u = 0; u < LK; u++)
            if (i0 + u < PER) {
                const int p = 32 * (i0 + u) + lane;
                // (PRE: the counts were looked up elsewhere -- routed to the shards that own the k-mers -- and wait at the
                // windows' positions; a window that was never sent holds 0.  A compile-time switch: the probing variant must
                // not carry the other one's control flow -- as a run-time branch it cost 1.5 KB of spills)
                unsigned v;
                if (PRE) v = p < nwin ? pre[p] : 0u;
                else v = lookup_settle(slots, geo, key[u], ok[u], q[u]).x;
                if (v < geo.floor) v = 0;          // a `dump -L floor` view: rarer k-mers are not in that table
                if (v < 1) v = 1;                  // fastaToKmerCoverageStats.cpp:328-330 (also windows with a non-base)
                const bool live = p < nwin;
                x[i0 + u] = live ? v : 0xFFFFFFFFu;
                if (live) { sw.cov[used + p] = v; mn = min(mn, v); mx = max(mx, v); part += v; }
            }
    }
    const unsigned lo = __reduce_min_sync(FULL, mn), hi = __reduce_max_sync(FULL, mx);
    unsigned long long sum;
    if (hi < (1u << 19)) {                         // 256 windows x 2^19 < 2^32: one redux instead of a 64-bit butterfly
        sum = __reduce_add_sync(FULL, (unsigned)part);
    } else {
        sum = part;
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    }
    mean = __fdiv_rn(__ll2float_rn((long long)sum), __ull2float_rn((unsigned long long)nwin));   // `long` sum, exact
    median = warp_median_regs<PER>(x, nwin, lo, hi);
}

// sequential fp32 sums of squares of a batch, one read per lane (fastaToKmerCoverageStats.cpp:389-402): strict read order,
// two roundings per term, no FMA (x86-64 -O2 without -march)
__device__ __forceinline__ void stats_flush(StatsWarp& sw, unsigned nb, float* __restrict__ stdev, int lane) {
    __syncwarp();
    if ((unsigned)lane < nb) {
        const unsigned n = sw.b_n[lane];
        const uint32_t* c = sw.cov + sw.b_off[lane];
        const float avg = sw.b_avg[lane];
        float sd;
        if (n == 1) {
            sd = __int_as_float(X86_DEFAULT_NAN_BITS);   // 0/0 on SSE = default NaN with the sign bit set ("-nan")
        } else {
            float acc = 0.0f;
            unsigned i = 0;
            for (; i + 4 <= n; i += 4) {
                const float d0 = __fsub_rn(__uint2float_rn(c[i]), avg), d1 = __fsub_rn(__uint2float_rn(c[i + 1]), avg);
                const float d2 = __fsub_rn(__uint2float_rn(c[i + 2]), avg), d3 = __fsub_rn(__uint2float_rn(c[i + 3]), avg);
                acc = __fadd_rn(acc, __fmul_rn(d0, d0));
                acc = __fadd_rn(acc, __fmul_rn(d1, d1));
                acc = __fadd_rn(acc, __fmul_rn(d2, d2));
                acc = __fadd_rn(acc, __fmul_rn(d3, d3));
            }
            for (; i < n; i++) {
                const float d = __fsub_rn(__uint2float_rn(c[i]), avg);
                acc = __fadd_rn(acc, __fmul_rn(d, d));
            }
            sd = __fsqrt_rn(__fdiv_rn(acc, __int2float_rn((int)n - 1)));
        }
        stdev[sw.b_r[lane]] = sd;
    }
    __syncwarp();
}

template <bool PRE>          // PRE: statistics from counts looked up elsewhere (routed lookups) instead of probing the table
__global__ void __launch_bounds__(PR_WARPS * 32, 4)
k_cov_stats(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, uint64_t nreads,
            int k, int canonical, const Slot* __restrict__ slots, Geo geo, uint32_t* __restrict__ median,
            float* __restrict__ mean, float* __restrict__ stdev, uint32_t* __restrict__ per_kmer, LongList ll,
            const uint32_t* __restrict__ order, int arena, const uint32_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    StatsWarp& sw = *reinterpret_cast<StatsWarp*>(dyn + (size_t)w * stats_warp_bytes(arena));
    const uint64_t i0 = ((uint64_t)blockIdx.x * PR_WARPS + w) * ST_RPW;      // position in processing order
    if (i0 >= nreads) return;
    const unsigned mk = kmask(k);
    // the warp's reads: lane rr holds the index and the two offsets of the rr-th one, handed round by shuffles
    static_assert(ST_RPW <= 32, "one read per lane");
    uint64_t my_r = 0, my_o0 = 0, my_o1 = 0;
    if (lane < ST_RPW && i0 + lane < nreads) {
        my_r = order ? (uint64_t)order[i0 + lane] : i0 + lane;
        my_o0 = offs[my_r]; my_o1 = offs[my_r + 1];
    }
    unsigned nb = 0, used = 0;
    for (int rr = 0; rr < ST_RPW && i0 + rr < nreads; rr++) {
        const uint64_t r = __shfl_sync(FULL, my_r, rr);
        const uint64_t o0 = __shfl_sync(FULL, my_o0, rr), o1 = __shfl_sync(FULL, my_o1, rr);
        const int L = (int)(o1 - o0 - 1);            // the record's last byte is its '\n' terminator
        const int nwin = L >= k ? L - k + 1 : 0;
        if (nwin > PR_MAXWIN) {
            if (lane == 0) {
                const unsigned slot = atomicAdd(ll.count, 1u);
                ll.idx[slot] = (unsigned)r;
                atomicMax(ll.max_win, (unsigned)nwin);
            }
            continue;
        }
        if (nwin == 0) {   // S6: shorter than k -> empty vector; S7-S9 on n = 0: 0, 0, sqrt(0/-1) = -0
            if (lane == 0) { median[r] = 0u; mean[r] = 0.0f; stdev[r] = __int_as_float(0x80000000); }
            continue;
        }
        if (nb == ST_BATCH || used + nwin > (unsigned)arena) { stats_flush(sw, nb, stdev, lane); nb = 0; used = 0; }
        const uint8_t* seq = recs + (o0 - rec_base);
        const uint32_t* pre = PRE ? counts + (o0 - rec_base) : nullptr;
        if (!PRE) front_planes(sw.f, seq, L, lane);
        uint32_t med; float mu;
        switch ((nwin + 31) >> 5) {
            case 1: stats_read<1, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            case 2: stats_read<2, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            case 3: stats_read<3, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            case 4: stats_read<4, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            case 5: stats_read<5, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            case 6: stats_read<6, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            case 7: stats_read<7, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
            default: stats_read<8, PRE>(sw, slots, geo, nwin, k, mk, canonical != 0, used, lane, med, mu, pre); break;
        }
        if (per_kmer) {
            __syncwarp();
            uint32_t* out = per_kmer + (o0 - rec_base);
            for (int p = lane; p < nwin; p += 32) out[p] = sw.cov[used + p];
        }
        if (lane == 0) {
            median[r] = med; mean[r] = mu;
            sw.b_off[nb] = used; sw.b_n[nb] = (unsigned)nwin; sw.b_r[nb] = (unsigned)r; sw.b_avg[nb] = mu;
        }
        nb++;
        used += (unsigned)nwin;
    }
    stats_flush(sw, nb, stdev, lane);
}

// k = 32: the warp-per-read kernels build their keys with the constant tag (k <= 31, tg_device.cuh), so every read takes
// the CTA-per-read path instead -- the long list is simply all reads.  (The caller has zeroed {count, max_win}.)
__global__ void __launch_bounds__(256)
k_list_all_reads(const uint64_t* __restrict__ offs, uint64_t nreads, int k, LongList ll) {
    unsigned mw = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nreads; i += (uint64_t)gridDim.x * blockDim.x) {
        ll.idx[i] = (unsigned)i;
        const long long L = (long long)(offs[i + 1] - offs[i]) - 1;
        if (L >= k) mw = max(mw, (unsigned)(L - k + 1));
    }
    mw = __reduce_max_sync(FULL, mw);
    if ((threadIdx.x & 31) == 0 && mw) atomicMax(ll.max_win, mw);
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(ll.count, (unsigned)nreads); atomicMax(ll.max_win, 1u); }   // (scratch sizes are never zero)
}
static cudaError_t launch_list_all_reads(const uint64_t* d_offs, uint64_t nreads, int k, LongList ll, cudaStream_t s) {
    uint64_t blocks = (nreads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_list_all_reads<<<(unsigned)blocks, 256, 0, s>>>(d_offs, nreads, k, ll);
    return cudaGetLastError();
}

cudaError_t launch_cov_stats(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                             int canonical, const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean,
                             float* d_stdev, uint32_t* d_per_kmer, LongList ll, const uint32_t* d_order, int arena,
                             const uint32_t* d_counts, cudaStream_t s) {
    TimedLaunch timed("k_cov_stats", s);
    if (nreads == 0) return cudaSuccess;
    if (k > 31) return launch_list_all_reads(d_offs, nreads, k, ll, s);
    if (arena < PR_MAXWIN) arena = PR_MAXWIN;              // one read of the warp path must fit
    const size_t dyn = stats_warp_bytes(arena) * PR_WARPS;
    const void* kern = d_counts ? (const void*)k_cov_stats<true> : (const void*)k_cov_stats<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    const uint64_t per_cta = (uint64_t)PR_WARPS * ST_RPW;
    const uint64_t blocks = (nreads + per_cta - 1) / per_cta;
    if (d_counts)
        k_cov_stats<true><<<(unsigned)blocks, PR_WARPS * 32, dyn, s>>>(d_recs, d_offs, rec_base, nreads, k, canonical, slots, geo,
                                                                       d_median, d_mean, d_stdev, d_per_kmer, ll, d_order, arena, d_counts);
    else
        k_cov_stats<false><<<(unsigned)blocks, PR_WARPS * 32, dyn, s>>>(d_recs, d_offs, rec_base, nreads, k, canonical, slots, geo,
                                                                        d_median, d_mean, d_stdev, d_per_kmer, ll, d_order, arena, d_counts);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// long path: one CTA per read (fastaToKmerCoverageStats.cpp:300-402)
// ---------------------------------------------------------------------------------------------------------
template <int GS>
__device__ __forceinline__ void read_cov_stats(const uint8_t* __restrict__ seq, int L, int k, int canonical,
                                               const Slot* __restrict__ slots, Geo geo, uint32_t* P0,
                                               uint32_t* P1, uint32_t* PB, uint32_t* cov, float* sq,
                                               unsigned long long* red, uint32_t* per_kmer, uint32_t& median,
                                               float& mean, float& stdev, int gtid, const uint32_t* __restrict__ pre = nullptr) {
    const int nwin = L >= k ? L - k + 1 : 0;
    if (nwin == 0) {   // S6: shorter than k -> empty vector; S7-S9 on n = 0: 0, 0, sqrt(0/-1) = -0
        median = 0; mean = 0.0f; stdev = __int_as_float(0x80000000);
        return;
    }
    const unsigned mk = kmask(k);
    const int nch = (L + 31) >> 5;
    if (!pre) pack_read_planes<GS>(seq, L, nch, P0, P1, PB, gtid);
    gsync<GS>();

    unsigned long long part = 0;
    for (int pb = 0; pb < nwin; pb += GS) {     // every lane runs every iteration: table_lookup is warp-convergent
        const int p = pb + gtid;
        const bool live = p < nwin;
        bool ok = false;
        unsigned long long key = 0ull;
        if (live && !pre) {
            const int c = p >> 5, o = p & 31;
            const unsigned bad = __funnelshift_r(PB[c], PB[c + 1], o) & mk;
            if (!bad) {
                const unsigned f0 = __funnelshift_r(P0[c], P0[c + 1], o) & mk;
                const unsigned f1 = __funnelshift_r(P1[c], P1[c + 1], o) & mk;
                key = make_key_k(f0, f1, k);
                if (canonical) {
                    const unsigned long long kr = make_key_k(rc_plane(f0, k), rc_plane(f1, k), k);
                    key = kr < key ? kr : key;
                }
                ok = true;
            }
        }
        unsigned v = pre ? (live ? pre[p] : 0u) : table_lookup(slots, geo, key, ok);
        if (live) {
            if (v < geo.floor) v = 0;              // a `dump -L floor` view
            if (v < 1) v = 1;                      // fastaToKmerCoverageStats.cpp:328-330
            cov[p] = v;
            if (per_kmer) per_kmer[p] = v;
            part += v;
        }
    }
    const unsigned long long sum = group_sum_u64<GS>(part, red, gtid);   // `long` sum, exact
    const float avg = __fdiv_rn(__ll2float_rn((long long)sum), __ull2float_rn((unsigned long long)nwin));
    gsync<GS>();
    for (int p = gtid; p < nwin; p += GS) {
        const float d = __fsub_rn(__uint2float_rn(cov[p]), avg);
        sq[p] = __fmul_rn(d, d);               // two roundings, no FMA (x86-64 -O2 without -march)
    }
    gsync<GS>();
    float sd;
    if (nwin == 1) {
        sd = __int_as_float(X86_DEFAULT_NAN_BITS);   // 0/0 on SSE = default NaN with the sign bit set ("-nan")
    } else {
        float acc = 0.0f;
        if (gtid == 0) {
            for (int p = 0; p < nwin; p++) acc = __fadd_rn(acc, sq[p]);     // strict read order
            acc = __fsqrt_rn(__fdiv_rn(acc, __int2float_rn(nwin - 1)));
        }
        sd = acc;
    }
    // median: odd -> middle, even -> u32 (wrapping) mean of the two middles
    const unsigned n2 = next_pow2((unsigned)nwin);
    for (unsigned p = nwin + gtid; p < n2; p += GS) cov[p] = 0xFFFFFFFFu;
    gsync<GS>();
    bitonic_sort<GS, uint32_t>(cov, n2, gtid);
    median = (nwin & 1) ? cov[nwin / 2] : (uint32_t)(cov[(nwin - 1) / 2] + cov[nwin / 2]) / 2u;
    mean = avg;
    stdev = sd;    // meaningful in gtid 0 only
}

// scratch layout per CTA of the long path: planes 3*(nch+1) u32 | cov n2 u32 | sq n2 f32
__host__ __device__ static inline size_t long_nch(unsigned max_win, int k) { return ((size_t)max_win + k - 1 + 31) / 32 + 2; }
__host__ __device__ static inline size_t pow2_ge(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
__host__ __device__ static inline size_t long_scratch_words(unsigned max_win, int k, int mult) {
    return 3 * long_nch(max_win, k) + (size_t)2 * pow2_ge((size_t)mult * max_win);
}

// The CTA-per-read kernels run in one of two ways.  Host-driven (the host-buffer entry points, which synchronise
// anyway): the host has read {count, max_win}, sized the scratch and passes them.  Device-driven (the *_dev entry
// points, which must not synchronise -- a host stall would leave the GPU idle): the kernel is launched unconditionally
// behind the warp-path kernel, reads {count, max_win} from `hdr` itself, lays the fixed scratch budget out and leaves
// at once when there is no long read.  A read too long for the budget raises error 4 instead of a wrong answer.
// all_reads != 0 (k-mers shorter than the fast path's 8): every read 0 .. n_long-1 goes through this kernel.
struct LongPlan { unsigned n_long, max_win, stride; size_t words_per_cta; };
__device__ __forceinline__ bool long_plan(const unsigned int* hdr, unsigned n_long, unsigned max_win, size_t words_per_cta,
                                          unsigned long long scratch_words, int k, int mult, int* error, LongPlan& pl) {
    pl.n_long = n_long; pl.max_win = max_win; pl.words_per_cta = words_per_cta; pl.stride = gridDim.x;
    if (!hdr) return true;
    pl.n_long = hdr[0]; pl.max_win = hdr[1];
    if (pl.n_long == 0) return false;
    pl.words_per_cta = long_scratch_words(pl.max_win, k, mult);
    const unsigned long long fit = scratch_words / pl.words_per_cta;
    if (fit == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(error, 4);
        return false;
    }
    pl.stride = (unsigned)min((unsigned long long)gridDim.x, fit);
    return blockIdx.x < pl.stride;
}
size_t cov_stats_long_scratch_bytes(unsigned max_win, int k, int nctas) {
    return long_scratch_words(max_win, k, 1) * 4 * (size_t)nctas;
}
size_t assign_long_scratch_bytes(unsigned max_win, int k, int nctas) {
    return long_scratch_words(max_win, k, 2) * 4 * (size_t)nctas;
}

__global__ void __launch_bounds__(LONG_THREADS)
k_cov_stats_long(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, int k,
                 int canonical, const Slot* __restrict__ slots, Geo geo, uint32_t* __restrict__ median,
                 float* __restrict__ mean, float* __restrict__ stdev, uint32_t* __restrict__ per_kmer,
                 const unsigned int* __restrict__ long_idx, unsigned int n_long_h, unsigned int max_win_h,
                 uint32_t* scratch, size_t words_per_cta_h, const unsigned int* __restrict__ hdr,
                 unsigned long long scratch_words, int* error, const uint32_t* __restrict__ counts) {
    __shared__ unsigned long long red[LONG_THREADS / 32];
    LongPlan pl;
    if (!long_plan(hdr, n_long_h, max_win_h, words_per_cta_h, scratch_words, k, 1, error, pl)) return;
    const unsigned n_long = pl.n_long, max_win = pl.max_win;
    const size_t nchw = ((size_t)max_win + k - 1 + 31) / 32 + 2;
    size_t n2max = 1; while (n2max < max_win) n2max <<= 1;
    uint32_t* base = scratch + (size_t)blockIdx.x * pl.words_per_cta;
    uint32_t* P0 = base; uint32_t* P1 = P0 + nchw; uint32_t* PB = P1 + nchw;
    uint32_t* cov = PB + nchw; float* sq = reinterpret_cast<float*>(cov + n2max);
    for (unsigned i = blockIdx.x; i < n_long; i += pl.stride) {
        const uint64_t r = long_idx ? long_idx[i] : i;
        const uint64_t o0 = offs[r], o1 = offs[r + 1];
        const int L = (int)(o1 - o0 - 1);
        uint32_t med; float mu, sd;
        read_cov_stats<LONG_THREADS>(recs + (o0 - rec_base), L, k, canonical, slots, geo, P0, P1, PB, cov, sq, red,
                                     per_kmer ? per_kmer + (o0 - rec_base) : nullptr, med, mu, sd, threadIdx.x,
                                     counts ? counts + (o0 - rec_base) : nullptr);
        if (threadIdx.x == 0) { median[r] = med; mean[r] = mu; stdev[r] = sd; }
        __syncthreads();
    }
}

cudaError_t launch_cov_stats_long(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k, int canonical,
                                  const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean, float* d_stdev,
                                  uint32_t* d_per_kmer, const unsigned int* d_long_idx, unsigned int n_long,
                                  unsigned int max_win, void* d_scratch, int nctas, cudaStream_t s) {
    TimedLaunch timed("k_cov_stats_long", s);
    if (n_long == 0) return cudaSuccess;
    k_cov_stats_long<<<nctas, LONG_THREADS, 0, s>>>(d_recs, d_offs, rec_base, k, canonical, slots, geo, d_median, d_mean,
                                                    d_stdev, d_per_kmer, d_long_idx, n_long, max_win,
                                                    (uint32_t*)d_scratch, long_scratch_words(max_win, k, 1), nullptr, 0,
                                                    nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_cov_stats_long_auto(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k,
                                       int canonical, const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean,
                                       float* d_stdev, uint32_t* d_per_kmer, LongList ll, void* d_scratch,
                                       size_t scratch_bytes, int* d_error, int nctas, cudaStream_t s, const uint32_t* d_counts) {
    TimedLaunch timed("k_cov_stats_long", s);
    k_cov_stats_long<<<nctas, LONG_THREADS, 0, s>>>(d_recs, d_offs, rec_base, k, canonical, slots, geo, d_median, d_mean,
                                                    d_stdev, d_per_kmer, ll.idx, 0, 0, (uint32_t*)d_scratch, 0, ll.count,
                                                    scratch_bytes / 4, d_error, d_counts);
    return cudaGetLastError();
}

// =========================================================================================================
// read -> bundle vote (ReadsToTranscripts.cc:216-274)
// =========================================================================================================
// entropy_ok is indexed [nG][nA][nT] (26^3 bytes); nC is implied for an all-ACGT window.
__device__ __forceinline__ bool window_entropy_ok(const uint8_t* __restrict__ lut, unsigned f0, unsigned f1, unsigned mk,
                                                  bool rc) {
    // codes: A=00 C=01 G=10 T=11 (bit1 = plane1, bit0 = plane0)
    const int nG = __popc(f1 & ~f0 & mk), nA = __popc(~f1 & ~f0 & mk), nT = __popc(f1 & f0 & mk);
    const int nC = __popc(~f1 & f0 & mk);
    // the reference evaluates the reverse-complemented string in the same G,A,T,C slot order:
    // its counts are (nC, nT, nA, nG) of the forward window
    return rc ? lut[(nC * 26 + nT) * 26 + nA] != 0 : lut[(nG * 26 + nA) * 26 + nT] != 0;
}

// labels of a settled lookup -> the reference's two lookups (forward window, reverse-complemented window):
// the slot of the canonical key holds the label of the bundle k-mer equal to the key (val) and of the bundle k-mer whose
// reverse complement is the key (aux).  A palindrome (even k only) is its own reverse complement.
__device__ __forceinline__ void labels_of(uint2 v, bool do_f, bool do_r, bool is_rc, bool pal, unsigned& vf, unsigned& vr) {
    vf = do_f ? (is_rc ? v.y : v.x) : 0u;
    vr = do_r ? ((is_rc || pal) ? v.x : v.y) : 0u;
}

// The reference sorts the hits and scans the runs (ReadsToTranscripts.cc:253-268): a label with m hits scores m-1, the
// largest label m-2, strict '>' while ascending => ties go to the smaller label.  The same result without a sort: walk
// the DISTINCT labels in ascending order (almost always one or two), one warp min and one warp count per label.
__device__ __forceinline__ void warp_vote(const int32_t* hits, int n, int lane, int& best, int& score) {
    int b = -1, sc = 0;
    if (n >= 2) {
        // nearly every read hits ONE bundle: one pass settles it (that label is also the largest: m - 2)
        const int first = hits[0];
        unsigned same = 0;
        int last = -1;
        for (int p = lane; p < n; p += 32) { const int h = hits[p]; same += h == first ? 1u : 0u; last = max(last, h); }
        same = __reduce_add_sync(FULL, same);
        if ((int)same == n) {
            sc = n - 2; b = first;
        } else {
            last = __reduce_max_sync(FULL, last);
            int cur = -1;
            while (true) {
                int mn = 0x7FFFFFFF;
                for (int p = lane; p < n; p += 32) { const int h = hits[p]; if (h > cur) mn = min(mn, h); }
                const int lab = __reduce_min_sync(FULL, mn);
                if (lab == 0x7FFFFFFF) break;
                unsigned m = 0;
                for (int p = lane; p < n; p += 32) m += hits[p] == lab ? 1u : 0u;
                m = __reduce_add_sync(FULL, m);
                const int s = (int)m - 1 - (lab == last ? 1 : 0);
                if (s > sc) { sc = s; b = lab; }
                cur = lab;
            }
        }
        if (sc <= 0) { b = -1; sc = 0; }
    }
    best = b; score = sc;
}

struct AssignWarp {
    WarpFront f;
    int32_t hits[2 * PR_MAXWIN];
};

// append the labels the warp's lanes hold (0 = none) to the hit list; nh is warp-uniform
__device__ __forceinline__ void push_hits(int32_t* hits, unsigned& nh, unsigned vf, unsigned vr, int lane) {
    const unsigned lt = (1u << lane) - 1u;
    const unsigned mf = __ballot_sync(FULL, vf != 0u), mr = __ballot_sync(FULL, vr != 0u);
    if (vf) hits[nh + __popc(mf & lt)] = (int32_t)vf - 1;
    nh += __popc(mf);
    if (vr) hits[nh + __popc(mr & lt)] = (int32_t)vr - 1;
    nh += __popc(mr);
}

template <int PER>
__device__ __forceinline__ void assign_read(AssignWarp& aw, const Slot* __restrict__ slots, const Geo& geo,
                                            const uint8_t* __restrict__ lut, int nwin, int k, unsigned mk, int strand, int lane,
                                            unsigned& nh_out) {
    constexpr int LK = 3;                          // rounds in flight
    unsigned nh = 0;
#pragma unroll
    for (int i0 = 0; i0 < PER; i0 += LK) {
        unsigned long long key[LK];
        unsigned fl[LK];                   // 1 = forward pass wanted, 2 = reverse pass wanted, 4 = key is the rc, 8 = palindrome
        bool ok[LK];
        LookupIssue q[LK];
#pragma unroll
        for (int u = 0; u < LK; u++)
            if (i0 + u < PER) {
                // label tables are keyed canonically whatever the library type: the strand flag only drops the second lookup
                const Window w = front_window(aw.f, 32 * (i0 + u) + lane, nwin, k, mk, true);
                bool do_f = false, do_r = false;
                if (w.valid) {             // a window with a non-ACGT character can never equal a table k-mer
                    do_f = window_entropy_ok(lut, w.f0, w.f1, mk, false);
                    do_r = !strand && window_entropy_ok(lut, w.f0, w.f1, mk, true);
                }
                // the probe does not wait for the entropy verdict (two LUT loads): nearly every window passes, and a window
                // that does not simply has its labels dropped when the probe is settled
                key[u] = w.key; ok[u] = w.valid;
                fl[u] = (do_f ? 1u : 0u) | (do_r ? 2u : 0u) | (w.is_rc ? 4u : 0u) | (w.pal ? 8u : 0u);
                q[u] = lookup_issue(slots, geo, key[u], ok[u]);
            }
#pragma unroll
        for (int u = 0; u < LK; u++)
            if (i0 + u < PER) {
                // ONE probe answers both passes of the reference (forward window, then reverse-complemented window)
                const uint2 v = lookup_settle(slots, geo, key[u], ok[u], q[u]);
                unsigned vf, vr2;
                labels_of(v, fl[u] & 1u, fl[u] & 2u, fl[u] & 4u, fl[u] & 8u, vf, vr2);
                push_hits(aw.hits, nh, vf, vr2, lane);
            }
    }
    __syncwarp();
    nh_out = nh;
}

// pct = (int)((float)max / num_kmer_pos * 100 + 0.5): fp32 divide, fp32 multiply, double add, truncate
__device__ __forceinline__ int assign_pct(int score, int nwin) {
    const float q = __fmul_rn(__fdiv_rn(__int2float_rn(score), __int2float_rn(nwin)), 100.0f);
    return (int)__dadd_rn((double)q, 0.5);
}

__global__ void __launch_bounds__(PR_WARPS * 32, 4)
k_assign(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, uint64_t nreads, int k,
         int strand, const Slot* __restrict__ slots, Geo geo, const uint8_t* __restrict__ lut,
         int32_t* __restrict__ best, int32_t* __restrict__ pct, int32_t* __restrict__ score, LongList ll,
         const uint32_t* __restrict__ order) {
    __shared__ AssignWarp smw[PR_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    AssignWarp& aw = smw[w];
    const uint64_t i = (uint64_t)blockIdx.x * PR_WARPS + w;
    if (i >= nreads) return;
    const uint64_t r = order ? (uint64_t)order[i] : i;
    const uint64_t o0 = offs[r], o1 = offs[r + 1];
    const int L = (int)(o1 - o0 - 1);
    const int nwin = L - k + 1;        // num_kmer_pos, may be <= 0
    if (nwin > PR_MAXWIN) {
        if (lane == 0) {
            const unsigned slot = atomicAdd(ll.count, 1u);
            ll.idx[slot] = (unsigned)r;
            atomicMax(ll.max_win, (unsigned)nwin);
        }
        return;
    }
    int b = -1, sc = 0, pc = 0;
    if (nwin > 0) {
        const unsigned mk = kmask(k);
        front_planes(aw.f, recs + (o0 - rec_base), L, lane);
        unsigned nh = 0;
        switch ((nwin + 31) >> 5) {
            case 1: assign_read<1>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            case 2: assign_read<2>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            case 3: assign_read<3>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            case 4: assign_read<4>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            case 5: assign_read<5>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            case 6: assign_read<6>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            case 7: assign_read<7>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
            default: assign_read<8>(aw, slots, geo, lut, nwin, k, mk, strand, lane, nh); break;
        }
        warp_vote(aw.hits, (int)nh, lane, b, sc);
        pc = assign_pct(sc, nwin);
    }
    if (lane == 0) { best[r] = b; pct[r] = pc; if (score) score[r] = sc; }
}

cudaError_t launch_assign(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                          int strand, const Slot* slots, Geo geo, const uint8_t* d_entropy_ok, int32_t* d_best,
                          int32_t* d_pct, int32_t* d_score, LongList ll, const uint32_t* d_order, cudaStream_t s) {
    TimedLaunch timed("k_assign", s);
    if (nreads == 0) return cudaSuccess;
    if (k > 31) return launch_list_all_reads(d_offs, nreads, k, ll, s);
    const uint64_t blocks = (nreads + PR_WARPS - 1) / PR_WARPS;
    k_assign<<<(unsigned)blocks, PR_WARPS * 32, 0, s>>>(d_recs, d_offs, rec_base, nreads, k, strand, slots, geo,
                                                        d_entropy_ok, d_best, d_pct, d_score, ll, d_order);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// long path: one CTA per read
// ---------------------------------------------------------------------------------------------------------
template <int GS>
__device__ __forceinline__ void read_assign(const uint8_t* __restrict__ seq, int L, int k, int strand,
                                            const Slot* __restrict__ slots, Geo geo,
                                            const uint8_t* __restrict__ lut, uint32_t* P0, uint32_t* P1, uint32_t* PB,
                                            int32_t* hits, unsigned int* nhits_p, int32_t& best, int32_t& score,
                                            int32_t& pct, int gtid) {
    const int nwin = L - k + 1;        // num_kmer_pos, may be <= 0
    best = -1; score = 0;
    if (nwin <= 0) { pct = 0; return; }
    const unsigned mk = kmask(k);
    const int nch = (L + 31) >> 5;
    if (gtid == 0) *nhits_p = 0;
    pack_read_planes<GS>(seq, L, nch, P0, P1, PB, gtid);
    gsync<GS>();
    for (int pb = 0; pb < nwin; pb += GS) {         // every lane runs every iteration: table_lookup2 is warp-convergent
        const int p = pb + gtid;
        bool do_f = false, do_r = false, is_rc = false, pal = false;
        unsigned long long key = 0ull;
        if (p < nwin) {
            const int c = p >> 5, o = p & 31;
            const unsigned bad = __funnelshift_r(PB[c], PB[c + 1], o) & mk;
            if (!bad) {                    // a window with a non-ACGT character can never equal a table k-mer
                const unsigned f0 = __funnelshift_r(P0[c], P0[c + 1], o) & mk;
                const unsigned f1 = __funnelshift_r(P1[c], P1[c + 1], o) & mk;
                do_f = window_entropy_ok(lut, f0, f1, mk, false);
                do_r = !strand && window_entropy_ok(lut, f0, f1, mk, true);
                const unsigned long long kf = make_key_k(f0, f1, k), kr = make_key_k(rc_plane(f0, k), rc_plane(f1, k), k);
                is_rc = kr < kf; pal = kr == kf;
                key = is_rc ? kr : kf;
            }
        }
        const uint2 v = table_lookup2(slots, geo, key, do_f || do_r);
        unsigned vf, vr;
        labels_of(v, do_f, do_r, is_rc, pal, vf, vr);
        if (vf) hits[atomicAdd(nhits_p, 1u)] = (int32_t)vf - 1;
        if (vr) hits[atomicAdd(nhits_p, 1u)] = (int32_t)vr - 1;
    }
    gsync<GS>();
    const int n = (int)*nhits_p;
    int b = -1, sc = 0;
    if (n >= 2) {
        const unsigned n2 = next_pow2((unsigned)n);
        for (unsigned p = n + gtid; p < n2; p += GS) hits[p] = 0x7FFFFFFF;
        gsync<GS>();
        bitonic_sort<GS, int32_t>(hits, n2, gtid);
        // a label with m hits scores m-1, the last (largest) label m-2; strict '>' while scanning ascending
        // labels => ties go to the smaller label (ReadsToTranscripts.cc:253-268)
        const int32_t last = hits[n - 1];
        for (int i = gtid; i < n; i += GS) {
            const int32_t h = hits[i];
            if (i == n - 1 || hits[i + 1] != h) {        // end of a run: multiplicity by lower_bound
                int lo = 0, hi = i;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (hits[mid] < h) lo = mid + 1; else hi = mid; }
                const int m = i + 1 - lo;
                const int s = m - 1 - (h == last ? 1 : 0);
                if (s > sc || (s == sc && s > 0 && h < b)) { sc = s; b = h; }
            }
        }
        // group arg-max (score desc, label asc)
        for (int o = 16; o > 0; o >>= 1) {
            const int os = __shfl_xor_sync(FULL, sc, o), ob = __shfl_xor_sync(FULL, b, o);
            if (os > sc || (os == sc && os > 0 && ob < b)) { sc = os; b = ob; }
        }
        __shared__ int wsc[LONG_THREADS / 32], wb[LONG_THREADS / 32];
        gsync<GS>();
        if ((gtid & 31) == 0) { wsc[gtid >> 5] = sc; wb[gtid >> 5] = b; }
        gsync<GS>();
        sc = wsc[0]; b = wb[0];
        for (int w = 1; w < GS / 32; w++)
            if (wsc[w] > sc || (wsc[w] == sc && wsc[w] > 0 && wb[w] < b)) { sc = wsc[w]; b = wb[w]; }
        gsync<GS>();
        if (sc <= 0) { b = -1; sc = 0; }
    }
    best = b; score = sc;
    pct = assign_pct(sc, nwin);
}

__global__ void __launch_bounds__(LONG_THREADS)
k_assign_long(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, int k, int strand,
              const Slot* __restrict__ slots, Geo geo, const uint8_t* __restrict__ lut, int32_t* __restrict__ best,
              int32_t* __restrict__ pct, int32_t* __restrict__ score, const unsigned int* __restrict__ long_idx,
              unsigned int n_long_h, unsigned int max_win_h, uint32_t* scratch, size_t words_per_cta_h,
              const unsigned int* __restrict__ hdr, unsigned long long scratch_words, int* error) {
    __shared__ unsigned int nhits;
    LongPlan pl;
    if (!long_plan(hdr, n_long_h, max_win_h, words_per_cta_h, scratch_words, k, 2, error, pl)) return;
    const unsigned n_long = pl.n_long, max_win = pl.max_win;
    const size_t nchw = ((size_t)max_win + k - 1 + 31) / 32 + 2;
    uint32_t* base = scratch + (size_t)blockIdx.x * pl.words_per_cta;
    uint32_t* P0 = base; uint32_t* P1 = P0 + nchw; uint32_t* PB = P1 + nchw;
    int32_t* hits = reinterpret_cast<int32_t*>(PB + nchw);
    for (unsigned i = blockIdx.x; i < n_long; i += pl.stride) {
        const uint64_t r = long_idx ? long_idx[i] : i;
        const uint64_t o0 = offs[r], o1 = offs[r + 1];
        const int L = (int)(o1 - o0 - 1);
        int32_t b, sc, pc;
        read_assign<LONG_THREADS>(recs + (o0 - rec_base), L, k, strand, slots, geo, lut, P0, P1, PB, hits, &nhits, b, sc,
                                  pc, threadIdx.x);
        if (threadIdx.x == 0) { best[r] = b; pct[r] = pc; if (score) score[r] = sc; }
        __syncthreads();
    }
}

cudaError_t launch_assign_long(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k, int strand,
                               const Slot* slots, Geo geo, const uint8_t* d_entropy_ok, int32_t* d_best,
                               int32_t* d_pct, int32_t* d_score, const unsigned int* d_long_idx, unsigned int n_long,
                               unsigned int max_win, void* d_scratch, int nctas, cudaStream_t s) {
    TimedLaunch timed("k_assign_long", s);
    if (n_long == 0) return cudaSuccess;
    k_assign_long<<<nctas, LONG_THREADS, 0, s>>>(d_recs, d_offs, rec_base, k, strand, slots, geo, d_entropy_ok, d_best,
                                                 d_pct, d_score, d_long_idx, n_long, max_win, (uint32_t*)d_scratch,
                                                 long_scratch_words(max_win, k, 2), nullptr, 0, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_assign_long_auto(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k, int strand,
                                    const Slot* slots, Geo geo, const uint8_t* d_entropy_ok, int32_t* d_best,
                                    int32_t* d_pct, int32_t* d_score, LongList ll, void* d_scratch, size_t scratch_bytes,
                                    int* d_error, int nctas, cudaStream_t s) {
    TimedLaunch timed("k_assign_long", s);
    k_assign_long<<<nctas, LONG_THREADS, 0, s>>>(d_recs, d_offs, rec_base, k, strand, slots, geo, d_entropy_ok, d_best,
                                                 d_pct, d_score, ll.idx, 0, 0, (uint32_t*)d_scratch, 0, ll.count,
                                                 scratch_bytes / 4, d_error);
    return cudaGetLastError();
}


// =========================================================================================================
// counting read by read (jellyfish count / KmerCounter::add_sequence) in locus order
// =========================================================================================================
// The flat-tile and log kernels (tg_kernels.cu) need no read structure.  This one does -- it visits the reads in LOCUS
// order, and that is the point: the reads that cover one stretch of a transcript are counted together, so the slots their
// k-mers share are L2-resident while they are being incremented, and the table needs no log, no partition replay and no
// second pass over the k-mers: one ld.cg + one RED.ADD per window, straight into the table.  Reads of any length: a warp
// walks a long read in segments of PR_MAXWIN windows.
constexpr int CR_RPW = 16;            // reads per warp
constexpr int CR_LK = 4;              // rounds (of 32 windows) whose home-slot loads are in flight together

__global__ void __launch_bounds__(PR_WARPS * 32, 4)
k_count_reads(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, uint64_t nreads,
              int k, int canonical, TableView t, const uint32_t* __restrict__ order) {
    __shared__ WarpFront fw[PR_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpFront& f = fw[w];
    const uint64_t i0 = ((uint64_t)blockIdx.x * PR_WARPS + w) * CR_RPW;
    if (i0 >= nreads) return;
    const unsigned mk = kmask(k);
    unsigned claimed = 0;
    uint64_t my_o0 = 0, my_o1 = 0;
    if (lane < CR_RPW && i0 + lane < nreads) {
        const uint64_t r = order ? (uint64_t)order[i0 + lane] : i0 + lane;
        my_o0 = offs[r]; my_o1 = offs[r + 1];
    }
    for (int rr = 0; rr < CR_RPW && i0 + rr < nreads; rr++) {
        const uint64_t o0 = __shfl_sync(FULL, my_o0, rr), o1 = __shfl_sync(FULL, my_o1, rr);
        const long long L = (long long)(o1 - o0) - 1;            // the record's last byte is its '\n' terminator
        const long long nwin = L >= k ? L - k + 1 : 0;
        for (long long seg = 0; seg < nwin; seg += PR_MAXWIN) {
            const int nseg = (int)min((long long)PR_MAXWIN, nwin - seg);
            __syncwarp();
            front_planes(f, recs + (o0 - rec_base) + seg, nseg + k - 1, lane);
            for (int i = 0; 32 * i < nseg; i += CR_LK) {
                unsigned long long key[CR_LK], cur[CR_LK];
                unsigned cnt[CR_LK];
                Probe pr[CR_LK];
#pragma unroll
                for (int u = 0; u < CR_LK; u++) {
                    const Window wd = front_window(f, 32 * (i + u) + lane, nseg, k, mk, canonical != 0);
                    key[u] = wd.valid ? wd.key : 0ull;
                    cnt[u] = 1u;
                    // a homopolymer run is one k-mer many times over: counted once per round, by its first lane
                    const bool homo = wd.valid && (wd.f0 == 0u || wd.f0 == mk) && (wd.f1 == 0u || wd.f1 == mk);
                    if (__any_sync(FULL, homo)) {
                        const unsigned code = (wd.f0 & 1u) | ((wd.f1 & 1u) << 1);
#pragma unroll
                        for (unsigned c = 0; c < 4; c++) {
                            const unsigned m = __ballot_sync(FULL, homo && code == c);
                            if (homo && code == c) {
                                if (lane == __ffs(m) - 1) cnt[u] = (unsigned)__popc(m); else key[u] = 0ull;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < CR_LK; u++) {
                    cur[u] = 0ull;
                    if (key[u] != 0ull) {
                        if (probe_home(t.g, key[u], pr[u])) cur[u] = __ldcg(&t.slots[pr[u].base + pr[u].off].key);
                        else { key[u] = 0ull; atomicExch(t.error, 2); }
                    }
                }
#pragma unroll
                for (int u = 0; u < CR_LK; u++)
                    if (key[u] != 0ull) {
                        Slot* sl = table_upsert_slot(t, key[u], pr[u], cur[u], claimed);
                        if (sl) atomicAdd(&sl->val, cnt[u]);
                    }
                __syncwarp();   // lanes leave the probe loops at different times: reconverge before the next rounds
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
    if (lane == 0 && claimed) atomicAdd(t.n_claimed, (unsigned long long)claimed);
}

cudaError_t launch_count_reads(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                               int canonical, TableView t, const uint32_t* d_order, cudaStream_t s) {
    TimedLaunch timed("k_count_reads", s);
    if (nreads == 0) return cudaSuccess;
    const uint64_t per_cta = (uint64_t)PR_WARPS * CR_RPW;
    k_count_reads<<<(unsigned)((nreads + per_cta - 1) / per_cta), PR_WARPS * 32, 0, s>>>(d_recs, d_offs, rec_base, nreads, k,
                                                                                        canonical, t, d_order);
    return cudaGetLastError();
}

// =========================================================================================================
// locus signature: the smallest strand-symmetric m-mer hash of a read (see LOCUS ORDER at the top of the file)
// =========================================================================================================
// m = min(k, 16): long enough to be (nearly) unique to its place in a transcriptome, short enough that a sequencing error
// rarely creates the read's smallest m-mer (an error touches m of the ~L m-mers of a read).  A read and its reverse
// complement get the same signature (both strands of an m-mer hash alike), so both mates' orientations cluster together.
//
// Flat scan like the count kernels (no read structure, except to know which read a window belongs to): a CTA takes 8 KiB
// of the record buffer per iteration, transposes it into bit planes (SWAR, four bases per lane), and every thread hashes
// the 32 m-mers that start in its 32-base chunk -- all shifts of two register pairs -- keeping a running minimum per
// record, which leaves through one atomicMin per (thread, record).  No padding is assumed behind the records: bytes at and
// past nbytes read as terminators.
constexpr int LC_GROUPS = CT_TILE / 128 + 1;      // 128-base SWAR groups per tile, + the halo group

__device__ __forceinline__ unsigned locus_hash(unsigned f0, unsigned f1, int m) {
    const unsigned r0 = __brev(~f0) >> (32 - m), r1 = __brev(~f1) >> (32 - m);
    const unsigned a = f0 * 0x9E3779B1u + f1 * 0x85EBCA77u, b = r0 * 0x9E3779B1u + r1 * 0x85EBCA77u;
    unsigned x = a < b ? a : b;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    return x ^ (x >> 13);
}

__global__ void __launch_bounds__(CT_THREADS, 4)
k_locus_tiles(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t nrec, uint64_t rec_base, int m,
              uint32_t* __restrict__ sig, uint32_t* __restrict__ idx) {
    // planes of the tile's chunks (+ halo group); pn = "is the record terminator": every record ends with exactly one '\n'
    // (the record-buffer format), so the record of a position is the record of the tile's first byte plus the terminators
    // before it -- ONE binary search per tile instead of one per thread
    __shared__ uint32_t p0[4 * LC_GROUPS], p1[4 * LC_GROUPS], pb[4 * LC_GROUPS], pn[4 * LC_GROUPS];
    __shared__ uint32_t wsum[CT_THREADS / 32];
    __shared__ unsigned long long tile_rec;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mm = kmask(m);
    const uint64_t nbytes = offs[nrec] - rec_base;
    const uint64_t ntiles = (nbytes + CT_TILE - 1) / CT_TILE;
    for (uint64_t i = blockIdx.x * (uint64_t)CT_THREADS + tid; i < nrec; i += (uint64_t)gridDim.x * CT_THREADS) idx[i] = (uint32_t)i;
    const unsigned* words = reinterpret_cast<const unsigned*>(recs);       // record buffers are at least 4-byte aligned
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t t0 = tile * CT_TILE;
        if (warp == 0) {                           // record of the tile's first byte: last offs[i] <= rec_base + t0
            // 32-ary search: every step the lanes probe 32 evenly spaced candidates of [lo, hi) at once, so the answer
            // takes ~5 dependent loads instead of ~24
            const uint64_t g0 = rec_base + t0;
            uint64_t lo = 0, hi = nrec;            // invariant: offs[lo] <= g0 < offs[hi]
            while (hi - lo > 1) {
                const uint64_t width = hi - lo;
                auto cand = [&](int j) { return lo + width * (uint64_t)(j + 1) / 33ull; };      // lo <= cand(j) < hi, ascending
                const int n = __popc(__ballot_sync(FULL, offs[cand(lane)] <= g0));               // true for a prefix of the lanes
                const uint64_t nlo = n ? cand(n - 1) : lo, nhi = n < 32 ? cand(n) : hi;
                lo = nlo; hi = nhi;
            }
            if (lane == 0) tile_rec = lo;
        }
        for (int g = warp; g < LC_GROUPS; g += CT_THREADS / 32) {
            const uint64_t first = t0 + 128ull * g + 4ull * lane;           // this lane's four bases
            unsigned w = 0x0A0A0A0Au;
            if (first < nbytes) {
                w = words[first >> 2];
                if (first + 4 > nbytes) {                                   // the word that straddles the end of the records
                    const unsigned keep = (unsigned)(nbytes - first) * 8u;
                    w = (w & ((1u << keep) - 1u)) | (0x0A0A0A0Au << keep);
                }
            }
            const unsigned cc = (w >> 1) ^ (w >> 2);
            const unsigned u = w & 0xDFDFDFDFu;
            const unsigned ok = swar_zero_bytes(u ^ 0x41414141u) | swar_zero_bytes(u ^ 0x43434343u) |
                                swar_zero_bytes(u ^ 0x47474747u) | swar_zero_bytes(u ^ 0x54545454u);
            const unsigned n0 = swar_gather(cc), n1 = swar_gather(cc >> 1), nb = ~swar_gather(ok >> 7) & 0xFu;
            const unsigned nn = first < nbytes ? swar_gather(swar_zero_bytes(w ^ 0x0A0A0A0Au) >> 7) : 0u;   // (padding is no terminator)
            const int pos = 4 * (lane & 7);
            unsigned b0 = n0 << pos, b1 = n1 << pos, bb = nb << pos, bn = nn << pos;
            or_over_8_lanes(b0, b1, bb);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) bn |= __shfl_xor_sync(FULL, bn, o);
            if ((lane & 7) == 0) { const int c = 4 * g + (lane >> 3); p0[c] = b0; p1[c] = b1; pb[c] = bb; pn[c] = bn; }
        }
        __syncthreads();
        const unsigned a0 = p0[tid], a1 = p1[tid], ab = pb[tid], an = pn[tid];
        const unsigned c0 = p0[tid + 1], c1 = p1[tid + 1], cb = pb[tid + 1];
        // terminators before this thread's chunk: exclusive scan of popc(an) over the CTA
        const unsigned mine = (unsigned)__popc(an);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned before = incl - mine;
        for (int w2 = 0; w2 < warp; w2++) before += wsum[w2];
        if (ab != FULL) {                          // (an m-mer STARTS in this chunk only if the chunk has a base)
            uint64_t rec = tile_rec + before;
            unsigned cur = 0xFFFFFFFFu;
            // Candidates: valid m-mers that start with A or C and end with G or T -- a quarter of them, and a property of
            // the m-mer that its reverse complement shares (rc swaps the two ends and complements them), so the sampled
            // set is the same on both strands and moves with the sequence, not with the read's start.  The thread walks
            // the set bits of its 32-position mask: ~8 hashes instead of 32.
            // bit s of ok_m: the m bases from position s on are all bases = no invalid flag in [s, s + m) of the 64-bit word
            // cb:ab, by OR-smearing (windows of 2, 4, .. p bits, then two overlapping p-windows cover m)
            unsigned long long inv = ((unsigned long long)cb << 32) | ab;
            int p = 1;
            while (2 * p <= m) { inv |= inv >> p; p *= 2; }
            inv |= inv >> (m - p);
            const unsigned ok_m = ~(unsigned)inv;
            const unsigned first_ac = ~a1;                                          // code A=00, C=01: high bit clear
            const unsigned last_gt = __funnelshift_r(a1, c1, (unsigned)(m - 1));     // code G=10, T=11: high bit set, at s + m - 1
            unsigned cand = ok_m & first_ac & last_gt;
            unsigned ends = an;                    // positions that END a record inside this chunk
            int done = 0;                          // positions below this have been accounted for record changes
            while (cand) {
                const int s = __ffs(cand) - 1;
                cand &= cand - 1u;
                // records that ended before position s: flush the running minimum once per record passed
                const unsigned passed = ends & ((1u << s) - 1u) & ~((1u << done) - 1u);
                if (passed) {
                    if (cur != 0xFFFFFFFFu) atomicMin(&sig[rec], cur);
                    cur = 0xFFFFFFFFu;
                    rec += (unsigned)__popc(passed);
                }
                done = s;
                cur = min(cur, locus_hash(__funnelshift_r(a0, c0, s) & mm, __funnelshift_r(a1, c1, s) & mm, m));
            }
            if (cur != 0xFFFFFFFFu) atomicMin(&sig[rec], cur);             // rec < nrec: a valid m-mer lies inside a record
        }
        __syncthreads();
    }
}

cudaError_t launch_read_locus(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int m,
                              uint32_t* d_sig, uint32_t* d_idx, int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_locus_tiles", s);
    if (nreads == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_sig, 0xFF, nreads * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_locus_tiles<<<sm_count * 4, CT_THREADS, 0, s>>>(d_recs, d_offs, nreads, rec_base, m, d_sig, d_idx);
    return cudaGetLastError();
}

}  // namespace tg
