// GraphFromFasta weldmer counting on the GPU.
//
// Reference: Chrysalis/analysis/GraphFromFasta.cc:1412-1424 sets a NonRedKmerTable up with the weldmer candidates (kk = 48
// bases by default, -kk) and streams every read through NonRedKmerTable::AddData (NonRedKmerTable.cc:162-200): the read is
// upper-cased, and every window of kk characters that EQUALS a stored weldmer (binary search over the sorted strings;
// forward strand only, no canonicalisation) bumps that weldmer's counter.  Weld decisions read the counters afterwards
// (GraphFromFasta.cc:538,584).
//
// Here the candidates sit in a small open-addressing table (built on the host: a few thousand to a few million 96-bit keys,
// L2-resident), and the reads are scanned by the same flat-tile machinery as the k-mer counter: TMA bulk copies of 8 KiB
// ASCII tiles, ballot transpose into bit planes, 32 windows per thread.  A kk-mer spans up to three 32-base plane words,
// so a window is three funnel shifts per plane.  Nearly every window misses (one 16-B load of an L2-resident slot).
#include "tg_internal.h"

namespace tg {

namespace {
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int WT_HALO = 64;                          // two extra chunks: a 48-base window starting at the tile's last base
constexpr int WT_LOAD = CT_TILE + WT_HALO;

struct WeldSmem {
    alignas(128) uint8_t ascii[2][WT_LOAD];
    uint32_t p0[CT_THREADS + 2], p1[CT_THREADS + 2], pb[CT_THREADS + 2];
    alignas(8) unsigned long long bar[2];
};
}

__host__ __device__ static inline unsigned long long weld_hash(unsigned long long lo, unsigned hi) {
    unsigned long long x = lo ^ ((unsigned long long)hi * 0x9E3779B97F4A7C15ull);
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
unsigned long long weld_hash_host(unsigned long long lo, unsigned hi) { return weld_hash(lo, hi); }

__global__ void __launch_bounds__(CT_THREADS, 3)
k_weld_tiles(const uint8_t* __restrict__ recs, uint64_t ntiles, int kk, WeldSlot* __restrict__ slots, uint64_t cap_mask) {
    __shared__ WeldSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hb = kk - 32;                          // bits of the window in its second plane word (1..16)
    const unsigned hm = (1u << hb) - 1u;
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) {
        mbar_arrive_expect_tx(&sm.bar[0], WT_LOAD);
        bulk_copy_g2s(sm.ascii[0], recs + tile * CT_TILE, WT_LOAD, &sm.bar[0]);
    }
    for (unsigned it = 0; tile < ntiles; it++, tile += gridDim.x) {
        const unsigned buf = it & 1u;
        const uint64_t next = tile + gridDim.x;
        if (tid == 0 && next < ntiles) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&sm.bar[buf ^ 1u], WT_LOAD);
            bulk_copy_g2s(sm.ascii[buf ^ 1u], recs + next * CT_TILE, WT_LOAD, &sm.bar[buf ^ 1u]);
        }
        mbar_wait(&sm.bar[buf], (it >> 1) & 1u);
        const uint8_t* a = sm.ascii[buf];
        for (int c = warp; c <= CT_THREADS + 1; c += CT_THREADS / 32) {
            const unsigned ch = a[c * 32 + lane];
            const unsigned code = base_code(ch);
            const unsigned b0 = __ballot_sync(FULL, code & 1u);
            const unsigned b1 = __ballot_sync(FULL, code >> 1);
            const unsigned bb = __ballot_sync(FULL, !base_valid(ch));
            if (lane == 0) { sm.p0[c] = b0; sm.p1[c] = b1; sm.pb[c] = bb; }
        }
        __syncthreads();   // planes complete; ascii[buf] is free for the TMA issued two iterations later

        const unsigned x0 = sm.p0[tid], y0 = sm.p0[tid + 1], z0 = sm.p0[tid + 2];
        const unsigned x1 = sm.p1[tid], y1 = sm.p1[tid + 1], z1 = sm.p1[tid + 2];
        const unsigned xb = sm.pb[tid], yb = sm.pb[tid + 1], zb = sm.pb[tid + 2];
        if (xb != FULL) {                            // (a window holds its first base: none starts in an all-invalid chunk)
#pragma unroll 4
            for (int s = 0; s < 32; s++) {
                // window = bits [s, s + kk) of the 96-bit strings x | y << 32 | z << 64
                const unsigned bad = __funnelshift_r(xb, yb, s) | (__funnelshift_r(yb, zb, s) & hm);
                if (bad) continue;
                const unsigned lo0 = __funnelshift_r(x0, y0, s), hi0 = __funnelshift_r(y0, z0, s) & hm;
                const unsigned lo1 = __funnelshift_r(x1, y1, s), hi1 = __funnelshift_r(y1, z1, s) & hm;
                const unsigned long long q0 = ((unsigned long long)hi0 << 32) | lo0, q1 = ((unsigned long long)hi1 << 32) | lo1;
                const unsigned long long klo = q0 | (q1 << 48);
                const unsigned khi = (unsigned)(q1 >> 16);
                uint64_t i = weld_hash(klo, khi) & cap_mask;
                for (uint64_t probes = 0; probes <= cap_mask; probes++) {
                    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(&slots[i]));
                    if (!(v.w & WELD_OCCUPIED)) break;                                   // free slot: not a weldmer
                    if ((((unsigned long long)v.y << 32) | v.x) == klo && v.z == khi) { atomicAdd(&slots[i].cnt, 1u); break; }
                    i = (i + 1) & cap_mask;
                }
            }
        }
        __syncthreads();   // planes consumed before the next iteration overwrites them
    }
}

cudaError_t launch_weld_tiles(const uint8_t* d_recs, uint64_t nbytes, int kk, WeldSlot* slots, uint64_t cap_mask, int sm_count,
                              cudaStream_t s) {
    TimedLaunch timed("k_weld_tiles", s);
    if (nbytes == 0) return cudaSuccess;
    const uint64_t ntiles = (nbytes + CT_TILE - 1) / CT_TILE;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k_weld_tiles, CT_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)per_sm * sm_count;
    if (grid > ntiles) grid = ntiles;
    k_weld_tiles<<<(unsigned)grid, CT_THREADS, 0, s>>>(d_recs, ntiles, kk, slots, cap_mask);
    return cudaGetLastError();
}

}  // namespace tg
