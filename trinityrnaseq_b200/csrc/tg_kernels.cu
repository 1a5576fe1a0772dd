// Hand-written sm_100a kernels of the Trinity k-mer hot path.
//
//   k_flat_tiles<COUNT>   jellyfish count / KmerCounter::add_sequence     
//   k_flat_tiles<LABEL>   ReadsToTranscripts bundle labelling             (§8a R3, R4)
//   k_load_pairs          fastaToKmerCoverageStats --kmers loader         (§8a S2)
//   k_cov_stats[_long]    compute_kmer_coverage / median / mean / stDev   (§8a S6-S9)
//   k_assign[_long]       ReadsToTranscripts per-read vote                (§8a R5, R7-R9)
//   k_histo, k_export     jellyfish histo / dump                          (§8a J2, J3)
//   k_gups                random-access roofline probes                   (§8d)
//
// None of this is a dense contraction: no tensor cores.  The flat-tile kernels stage ASCII read tiles into
// shared memory with TMA bulk copies (cp.async.bulk + mbarrier, double buffered), transpose them into
// bit planes with warp ballots, and then every thread rolls 32 windows out of two plane words.
#include "tg_internal.h"

namespace tg {

constexpr unsigned FULL = 0xFFFFFFFFu;

thread_local KernelTimer* g_kernel_timer = nullptr;

int max_resident_ctas(const void* kernel, int threads, size_t dyn_smem, int device) {
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem);
    if (device < 0) cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (per_sm < 1) per_sm = 1;
    return per_sm * sms;
}

// =========================================================================================================
// Flat tiles: count / label
// =========================================================================================================
enum { MODE_COUNT = 0, MODE_LABEL = 1 };

struct FlatSmem {
    alignas(128) uint8_t ascii[2][CT_LOAD];
    uint32_t p0[CT_THREADS + 1];
    uint32_t p1[CT_THREADS + 1];
    uint32_t pb[CT_THREADS + 1];
    alignas(8) unsigned long long bar[2];
};

template <int MODE>
__global__ void __launch_bounds__(CT_THREADS, MODE == 0 ? 4 : 3)    // LABEL carries labels + orientation bits: no spills at 3
k_flat_tiles(const uint8_t* __restrict__ recs, uint64_t ntiles, int k, int canonical, TableView t,
             const uint64_t* __restrict__ offs, uint64_t nrec, uint32_t first_index, uint64_t rec_base) {
    __shared__ FlatSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mk = kmask(k);
    unsigned claimed = 0;

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) {
        mbar_arrive_expect_tx(&sm.bar[0], CT_LOAD);
        bulk_copy_g2s(sm.ascii[0], recs + tile * CT_TILE, CT_LOAD, &sm.bar[0]);
    }

    for (unsigned it = 0; tile < ntiles; it++, tile += gridDim.x) {
        const unsigned buf = it & 1u;
        // prefetch the next tile into the other buffer (its previous contents were consumed before the
        // __syncthreads that ended the previous iteration's pack step)
        const uint64_t next = tile + gridDim.x;
        if (tid == 0 && next < ntiles) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&sm.bar[buf ^ 1u], CT_LOAD);
            bulk_copy_g2s(sm.ascii[buf ^ 1u], recs + next * CT_TILE, CT_LOAD, &sm.bar[buf ^ 1u]);
        }
        mbar_wait(&sm.bar[buf], (it >> 1) & 1u);

        // ---- ASCII -> bit planes: one ballot triple per 32-base chunk
        const uint8_t* a = sm.ascii[buf];
        for (int c = warp; c <= CT_THREADS; c += CT_THREADS / 32) {
            const unsigned ch = a[c * 32 + lane];
            const unsigned code = base_code(ch);
            const unsigned b0 = __ballot_sync(FULL, code & 1u);
            const unsigned b1 = __ballot_sync(FULL, code >> 1);
            const unsigned bb = __ballot_sync(FULL, !base_valid(ch));
            if (lane == 0) { sm.p0[c] = b0; sm.p1[c] = b1; sm.pb[c] = bb; }
        }
        __syncthreads();   // planes complete; ascii[buf] is free for the TMA issued two iterations later

        // ---- 32 windows per thread
        const unsigned a0 = sm.p0[tid], a1 = sm.p1[tid], ab = sm.pb[tid];
        const unsigned c0 = sm.p0[tid + 1], c1 = sm.p1[tid + 1], cb = sm.pb[tid + 1];

        uint32_t label = 0;
        uint64_t next_off = ~0ull;
        const uint64_t g0 = rec_base + tile * CT_TILE + (uint64_t)tid * 32;   // in the offs[] frame
        if (MODE == MODE_LABEL) {
            // record index of this thread's first base: last offs[i] <= g0
            uint64_t lo = 0, hi = nrec;   // candidates 0..nrec-1 (offs has nrec+1 entries), so lo+1 <= nrec
            while (hi - lo > 1) {
                uint64_t mid = (lo + hi) >> 1;
                if (offs[mid] <= g0) lo = mid; else hi = mid;
            }
            label = (uint32_t)lo;
            next_off = offs[lo + 1];
        }

        const bool has_windows = ab != FULL || cb != FULL;
        const unsigned act = __ballot_sync(FULL, has_windows);   // the lanes that enter the loop below together
        if (has_windows) {
#pragma unroll 1
            for (int g = 0; g < 32; g += 4) {
                unsigned long long key[4];
                Probe pr[4];
                unsigned long long cur[4];
                unsigned cnt[4];
                uint32_t lab[4];
                unsigned rcs = 0;      // LABEL: bit u set = window u is the reverse complement of its (canonical) key
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int s = g + u;
                    const unsigned bad = __funnelshift_r(ab, cb, s) & mk;
                    const unsigned f0 = __funnelshift_r(a0, c0, s) & mk;
                    const unsigned f1 = __funnelshift_r(a1, c1, s) & mk;
                    unsigned long long kf = make_key(f0, f1);
                    if (MODE == MODE_COUNT) {
                        if (canonical) {
                            const unsigned long long kr = make_key(rc_plane(f0, k), rc_plane(f1, k));
                            kf = kr < kf ? kr : kf;
                        }
                    } else {
                        // a valid window lies inside one record, so label < nrec whenever we advance
                        if (!bad) while (g0 + s >= next_off) { label++; next_off = offs[label + 1]; }
                        lab[u] = first_index + label + 1;
                        // label tables are keyed canonically with one label per orientation (tg_device.cuh)
                        const unsigned long long kr = make_key(rc_plane(f0, k), rc_plane(f1, k));
                        if (kr < kf) { kf = kr; rcs |= 1u << u; }
                    }
                    key[u] = bad ? 0ull : kf;
                    cnt[u] = 1;
                }
                if (MODE == MODE_COUNT) {
                    // run-length merge inside the group: homopolymer runs hit one slot once, not four times
#pragma unroll
                    for (int u = 1; u < 4; u++)
                        if (key[u] != 0ull && key[u] == key[u - 1]) { cnt[u] += cnt[u - 1]; key[u - 1] = 0ull; }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (key[u] != 0ull) {
                        if (probe_home(t.g, key[u], pr[u])) cur[u] = __ldcg(&t.slots[pr[u].base + pr[u].off].key);
                        else { key[u] = 0ull; atomicExch(t.error, 2); }
                    }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (key[u] != 0ull) {
                        if (MODE == MODE_COUNT) {
                            Slot* sl = table_upsert_slot(t, key[u], pr[u], cur[u], claimed);
                            if (sl) atomicAdd(&sl->val, cnt[u]);
                        } else {
                            // `claimed` counts distinct FORWARD k-mers (NonRedKmerTable's size): first label of a field
                            unsigned slots_claimed = 0;
                            Slot* sl = table_upsert_slot(t, key[u], pr[u], cur[u], slots_claimed);
                            if (sl && atomicMax((rcs >> u) & 1u ? &sl->aux : &sl->val, lab[u]) == 0u) claimed++;
                        }
                    }
                __syncwarp(act);   // lanes leave the probe loops at different times: reconverge before the next group
            }
        }
        __syncthreads();   // planes consumed before the next iteration overwrites them
    }

    // distinct-key accounting: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
    if (lane == 0 && claimed) atomicAdd(t.n_claimed, (unsigned long long)claimed);
}

static int flat_grid(const void* kern, uint64_t ntiles, int sm_count) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CT_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    uint64_t g = (uint64_t)per_sm * sm_count;
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_count_tiles(const uint8_t* d_recs, uint64_t nbytes, int k, int canonical, TableView t,
                               int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_flat_tiles<COUNT>", s);
    if (nbytes == 0) return cudaSuccess;
    const uint64_t ntiles = (nbytes + CT_TILE - 1) / CT_TILE;
    const int grid = flat_grid((const void*)k_flat_tiles<MODE_COUNT>, ntiles, sm_count);
    k_flat_tiles<MODE_COUNT><<<grid, CT_THREADS, 0, s>>>(d_recs, ntiles, k, canonical, t, nullptr, 0, 0, 0);
    return cudaGetLastError();
}

cudaError_t launch_label_tiles(const uint8_t* d_recs, uint64_t nbytes, const uint64_t* d_offs, uint64_t rec_base,
                               uint64_t nbundles, uint32_t first_bundle_index, int k, TableView t, int sm_count,
                               cudaStream_t s) {
    TimedLaunch timed("k_flat_tiles<LABEL>", s);
    if (nbytes == 0 || nbundles == 0) return cudaSuccess;
    const uint64_t ntiles = (nbytes + CT_TILE - 1) / CT_TILE;
    const int grid = flat_grid((const void*)k_flat_tiles<MODE_LABEL>, ntiles, sm_count);
    k_flat_tiles<MODE_LABEL><<<grid, CT_THREADS, 0, s>>>(d_recs, ntiles, k, 0, t, d_offs, nbundles, first_bundle_index,
                                                    rec_base);
    return cudaGetLastError();
}

// =========================================================================================================
// Partitioned count path.
//   phase 1  k_log_tiles   reads -> canonical k-mer keys -> appended to the log bin of their hash partition
//   phase 2  k_log_replay  bins replayed in order: all CAS/RED traffic of a bin lands in one L2-resident group of
//                          table partitions (prefetched into L2 with a bulk prefetch one bin ahead)
// The log costs 8 B written + 8 B read per occurrence, all of it sequential, instead of a random DRAM
// read-modify-write per occurrence.  Multi-GPU counting uses the same two kernels with an all-to-all of the
// bins between them (bin = global partition; rank r owns a contiguous range of bins).
// =========================================================================================================
// Phase 1 per 4 KiB tile of reads (256 threads, 16 windows each):
//   A  every valid window: key -> bin = partition of its hash; rank = shared atomicAdd on the bin's 16-bit counter
//      (two counters per word); (bin, rank) parked in shared memory
//   S  exclusive scan of the counters (-> place of each bin's run in the sorted tile) and ONE global atomicAdd per
//      non-empty bin reserving the run's place in the log -- all bins in parallel, one round trip per tile
//   B  keys recomputed (cheap: four funnel shifts, two bit reversals) and scattered into the sorted tile
//   W  the sorted tile streamed out: consecutive threads write consecutive entries of a run, so the stores of a
//      warp cover a handful of sectors instead of 32
// Homopolymer windows never enter the log (they would all hit one slot): they are tallied in registers and
// leave through lg.hpoly as four (key, count) pairs.
constexpr int LT_THREADS = 256;
constexpr int LT_WIN = 16;                          // windows per thread
constexpr int LT_TILE = LT_THREADS * LT_WIN;        // 4096 bases per tile
constexpr int LT_LOAD = LT_TILE + CT_HALO;
constexpr int LT_CHUNKS = LT_TILE / 32;             // 32-base plane words per tile
static_assert(CT_TILE % LT_TILE == 0, "record buffers are padded to CT_TILE");

struct LogSmem {
    alignas(128) uint8_t ascii[2][LT_LOAD];
    uint32_t p0[LT_CHUNKS + 1];
    uint32_t p1[LT_CHUNKS + 1];
    uint32_t pb[LT_CHUNKS + 1];
    uint32_t meta[LT_TILE];                         // bin << 12 | rank, ~0 = no entry
    uint32_t wtot[LT_THREADS / 32];
    uint32_t total;
    alignas(8) unsigned long long bar[2];
    unsigned long long* seg[LOG_MAX_RANKS];         // this rank's segment in every owner's log (dynamic index: not params)
};

__device__ __forceinline__ unsigned long long window_key(unsigned f0, unsigned f1, int k, int canonical) {
    unsigned long long key = make_key(f0, f1);
    if (canonical) {
        const unsigned long long kr = make_key(rc_plane(f0, k), rc_plane(f1, k));
        key = kr < key ? kr : key;
    }
    return key;
}

__global__ void __launch_bounds__(LT_THREADS, 2)
k_log_tiles(const uint8_t* __restrict__ recs, uint64_t ntiles, int k, int canonical, LogView lg, TableView t) {
    __shared__ LogSmem sm;
    extern __shared__ __align__(16) unsigned char dyn[];
    // dynamic: skey[LT_TILE] u64 | delta[nbins] u32 | sbin[LT_TILE] u16 | cnt16[nbins2] u16 | off16[nbins2] u16
    const unsigned nbins = lg.nbins, nbins2 = (nbins + 1u) & ~1u;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(dyn);
    unsigned int* delta = reinterpret_cast<unsigned int*>(skey + LT_TILE);
    unsigned short* sbin = reinterpret_cast<unsigned short*>(delta + nbins);
    unsigned short* cnt16 = sbin + LT_TILE;
    unsigned short* off16 = cnt16 + nbins2;
    unsigned int* cnt32 = reinterpret_cast<unsigned int*>(cnt16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mk = kmask(k);
    unsigned claimed = 0;
    unsigned hpA = 0, hpC = 0, hpG = 0, hpT = 0;    // homopolymer windows seen by this thread, by base code

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_mbar_init();
#pragma unroll
        for (int r = 0; r < LOG_MAX_RANKS; r++)
            sm.seg[r] = lg.owner[r] ? lg.owner[r] + (((unsigned long long)lg.src << lg.lp_shift) * lg.cap) : nullptr;
    }
    for (unsigned b = tid; b < nbins2 / 2; b += LT_THREADS) cnt32[b] = 0u;
    __syncthreads();

    uint64_t tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) {
        mbar_arrive_expect_tx(&sm.bar[0], LT_LOAD);
        bulk_copy_g2s(sm.ascii[0], recs + tile * LT_TILE, LT_LOAD, &sm.bar[0]);
    }
    const unsigned per = (nbins + LT_THREADS - 1) / LT_THREADS;
    for (unsigned it = 0; tile < ntiles; it++, tile += gridDim.x) {
        const unsigned buf = it & 1u;
        const uint64_t next = tile + gridDim.x;
        if (tid == 0 && next < ntiles) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&sm.bar[buf ^ 1u], LT_LOAD);
            bulk_copy_g2s(sm.ascii[buf ^ 1u], recs + next * LT_TILE, LT_LOAD, &sm.bar[buf ^ 1u]);
        }
        mbar_wait(&sm.bar[buf], (it >> 1) & 1u);
        const uint8_t* a = sm.ascii[buf];
        for (int c = warp; c <= LT_CHUNKS; c += LT_THREADS / 32) {
            const unsigned ch = a[c * 32 + lane];
            const unsigned code = base_code(ch);
            const unsigned b0 = __ballot_sync(FULL, code & 1u);
            const unsigned b1 = __ballot_sync(FULL, code >> 1);
            const unsigned bb = __ballot_sync(FULL, !base_valid(ch));
            if (lane == 0) { sm.p0[c] = b0; sm.p1[c] = b1; sm.pb[c] = bb; }
        }
        __syncthreads();   // planes complete; ascii[buf] is free for the TMA issued two iterations later

        // ---- A: bins and ranks
        const int c = tid >> 1, sh0 = (tid & 1) * LT_WIN;
        const unsigned a0 = sm.p0[c], a1 = sm.p1[c], ab = sm.pb[c];
        const unsigned c0 = sm.p0[c + 1], c1 = sm.p1[c + 1], cb = sm.pb[c + 1];
#pragma unroll 4
        for (int j = 0; j < LT_WIN; j++) {
            const int s = sh0 + j;
            const unsigned bad = __funnelshift_r(ab, cb, s) & mk;
            const unsigned f0 = __funnelshift_r(a0, c0, s) & mk;
            const unsigned f1 = __funnelshift_r(a1, c1, s) & mk;
            const bool homo = (f0 == 0u || f0 == mk) && (f1 == 0u || f1 == mk);
            unsigned m = 0xFFFFFFFFu;
            if (!bad) {
                if (homo) {
                    const unsigned code = (f0 & 1u) | ((f1 & 1u) << 1);
                    hpA += code == 0u; hpC += code == 1u; hpG += code == 2u; hpT += code == 3u;
                } else {
                    const unsigned bin = hash_part(mix64(window_key(f0, f1, k, canonical)), nbins);
                    const unsigned shift = (bin & 1u) * 16u;
                    const unsigned old = atomicAdd(&cnt32[bin >> 1], 1u << shift);
                    m = (bin << 12) | ((old >> shift) & 0xFFFFu);
                }
            }
            sm.meta[j * LT_THREADS + tid] = m;
        }
        __syncthreads();

        // ---- S: scan the bin counters, reserve every bin's run in the log
        {
            const unsigned b0 = tid * per, b1 = min(b0 + per, nbins);
            unsigned sum = 0;
            for (unsigned b = b0; b < b1; b++) sum += cnt16[b];
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) sm.wtot[warp] = incl;
            __syncthreads();
            unsigned run = incl - sum;
            for (int w = 0; w < warp; w++) run += sm.wtot[w];
            if (tid == LT_THREADS - 1) sm.total = run + sum;
            for (unsigned b = b0; b < b1; b++) {
                const unsigned n = cnt16[b];
                off16[b] = (unsigned short)run;
                if (n) {
                    unsigned base = *reinterpret_cast<volatile unsigned int*>(&lg.cursor[b]);
                    if (base < lg.cap) base = atomicAdd(&lg.cursor[b], n);   // a full bin is never advanced again
                    delta[b] = base - run;
                    cnt16[b] = 0;
                    run += n;
                }
            }
        }
        __syncthreads();

        // ---- B: keys again, into their sorted places
#pragma unroll 4
        for (int j = 0; j < LT_WIN; j++) {
            const unsigned m = sm.meta[j * LT_THREADS + tid];
            if (m != 0xFFFFFFFFu) {
                const int s = sh0 + j;
                const unsigned f0 = __funnelshift_r(a0, c0, s) & mk;
                const unsigned f1 = __funnelshift_r(a1, c1, s) & mk;
                const unsigned bin = m >> 12;
                const unsigned idx = off16[bin] + (m & 0xFFFu);
                skey[idx] = window_key(f0, f1, k, canonical);
                sbin[idx] = (unsigned short)bin;
            }
        }
        __syncthreads();

        // ---- W: stream the sorted tile out
        const unsigned total = sm.total;
        for (unsigned i = tid; i < total; i += LT_THREADS) {
            const unsigned bin = sbin[i];
            const unsigned pos = delta[bin] + i;
            const unsigned long long key = skey[i];
            if (pos < lg.cap) {
                // bin -> (owner, bin inside the owner); the store lands in local HBM or, over NVLink, in the owner's log
                const unsigned o = bin >> lg.lp_shift, lb = bin - (o << lg.lp_shift);
                sm.seg[o][(unsigned long long)lb * lg.cap + pos] = key;
            } else if (t.slots) {      // bin full: count this occurrence directly
                table_update<false>(t, key, 1u, claimed);
            } else {
                atomicExch(lg.error, 3);
            }
        }
        __syncthreads();   // everything consumed before the next tile reuses it
    }

    // homopolymer tallies: hpoly[c] = key of the window made of base code c, hpoly[4 + c] += occurrences
#pragma unroll
    for (int c = 0; c < 4; c++) {
        unsigned v = c == 0 ? hpA : c == 1 ? hpC : c == 2 ? hpG : hpT;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0 && v) {
            const unsigned f0 = (c & 1) ? mk : 0u, f1 = (c & 2) ? mk : 0u;
            lg.hpoly[c] = window_key(f0, f1, k, canonical);
            atomicAdd(&lg.hpoly[4 + c], (unsigned long long)v);
        }
    }
    if (t.slots) {
        for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
        if (lane == 0 && claimed) atomicAdd(t.n_claimed, (unsigned long long)claimed);
    }
}

size_t log_tiles_smem_bytes(unsigned nbins) {
    const unsigned nbins2 = (nbins + 1u) & ~1u;
    return (size_t)LT_TILE * 8 + (size_t)nbins * 4 + (size_t)LT_TILE * 2 + (size_t)nbins2 * 2 * 2;
}

cudaError_t launch_log_tiles(const uint8_t* d_recs, uint64_t nbytes, int k, int canonical, LogView lg, TableView t,
                             int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_log_tiles", s);
    if (nbytes == 0) return cudaSuccess;
    if (lg.nbins > LOG_MAX_BINS) return cudaErrorInvalidValue;
    const uint64_t ntiles = (nbytes + LT_TILE - 1) / LT_TILE;
    const size_t dyn = log_tiles_smem_bytes(lg.nbins);
    cudaError_t e = cudaFuncSetAttribute((const void*)k_log_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k_log_tiles, LT_THREADS, dyn);
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)per_sm * sm_count;
    if (grid > ntiles) grid = ntiles;
    k_log_tiles<<<(unsigned)grid, LT_THREADS, dyn, s>>>(d_recs, ntiles, k, canonical, lg, t);
    return cudaGetLastError();
}

// ---- phase 2 ------------------------------------------------------------------------------------------------
constexpr int RP_THREADS = 256;
constexpr int RP_PER_THREAD = 8;
constexpr int RP_CHUNK = RP_THREADS * RP_PER_THREAD;
#ifndef RP_GROUP
#define RP_GROUP 4                                  // entries per thread whose probes are in flight together
#endif
#ifndef RP_MIN_CTAS
#define RP_MIN_CTAS 4
#endif
static_assert(RP_PER_THREAD % RP_GROUP == 0, "groups tile a thread's entries");
constexpr int RP_FOLD = RP_CHUNK;                   // slots of the per-chunk fold table; entries that do not find a
constexpr int RP_FOLD_PROBES = 8;                   // place within a few probes go to the table directly
static_assert((RP_FOLD & (RP_FOLD - 1)) == 0 && (RP_FOLD / RP_THREADS) % 4 == 0, "fold table geometry");

// Replay order.  Bins are dealt round-robin to G GROUPS (bin lp belongs to group lp % G); the chunk index space
// lists group 0's bins first (in bin order, each bin's nsrc source segments together), then group 1's, ...  Every
// group hands its chunks out in order from its own counter, and the CTAs are split evenly over the groups, so at any
// time about G bins are being replayed, each by 1/G of the machine:
//   * G = 1 keeps the table traffic of a whole bin inside one L2-resident partition group, but every occurrence of
//     a hot k-mer then arrives within the few tens of microseconds its bin is open, and same-address atomics
//     serialise in L2;
//   * a static stride over all chunks (no counters) lets CTAs drift tens of bins apart: no hot spots, but the open
//     partitions no longer fit in L2 (measured: 83 GB of DRAM reads for a 5.5 GB table);
//   * G groups bound both: G partitions open, hot k-mers spread over G times the time.
// Layout of the plan array: chunk_start[nperm + 1] (prefix over the permuted segments), then G work counters.
__host__ __device__ __forceinline__ unsigned replay_per_group(unsigned nlocal, unsigned G) { return (nlocal + G - 1) / G; }

// permuted segment index pq -> (lp, src); lp >= nlocal means "no such bin" (padding of the last groups)
__device__ __forceinline__ void replay_segment(unsigned pq, unsigned nsrc, unsigned nlocal, unsigned G, unsigned& lp,
                                               unsigned& src) {
    const unsigned per = replay_per_group(nlocal, G);
    const unsigned bp = pq / nsrc;
    src = pq % nsrc;
    lp = (bp % per) * G + bp / per;
}

__global__ void __launch_bounds__(1024)
k_log_plan(const unsigned int* __restrict__ cursor, unsigned cap, unsigned nsrc, unsigned nlocal, unsigned G,
           unsigned long long* __restrict__ chunk_start) {
    __shared__ unsigned long long part[1024];
    const unsigned nperm = replay_per_group(nlocal, G) * G * nsrc;
    const unsigned per = (nperm + 1023) / 1024;
    const unsigned q0 = threadIdx.x * per, q1 = min(q0 + per, nperm);
    auto chunks_of = [&](unsigned pq) -> unsigned long long {
        unsigned lp, src;
        replay_segment(pq, nsrc, nlocal, G, lp, src);
        if (lp >= nlocal) return 0ull;
        const unsigned c = min(cursor[src * nlocal + lp], cap);
        return (c + RP_CHUNK - 1) / RP_CHUNK;
    };
    unsigned long long sum = 0;
    for (unsigned q = q0; q < q1; q++) sum += chunks_of(q);
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned long long v = threadIdx.x >= o ? part[threadIdx.x - o] : 0ull;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - sum;    // exclusive prefix
    for (unsigned q = q0; q < q1; q++) {
        chunk_start[q] = run;
        run += chunks_of(q);
    }
    if (threadIdx.x == 1023) chunk_start[nperm] = part[1023];
}

// the counters need values written by other threads: a second tiny kernel keeps k_log_plan simple
__global__ void k_log_plan_counters(unsigned nsrc, unsigned nlocal, unsigned G, unsigned long long* chunk_start) {
    const unsigned nperm = replay_per_group(nlocal, G) * G * nsrc;
    if (threadIdx.x < G) chunk_start[nperm + 1 + threadIdx.x] = chunk_start[threadIdx.x * replay_per_group(nlocal, G) * nsrc];
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <bool FOLD>
__global__ void __launch_bounds__(RP_THREADS, RP_MIN_CTAS)
k_log_replay(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ cursor, unsigned cap,
             unsigned nsrc, unsigned nlocal, unsigned bin0, unsigned nbins_global, unsigned G,
             unsigned long long* chunk_start, unsigned long long* hpoly, TableView t, int prefetch) {
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned per = replay_per_group(nlocal, G);
    const unsigned nperm = per * G * nsrc;
    unsigned claimed = 0;
    if (hpoly && blockIdx.x == 0 && tid < 4) {
        // homopolymer tallies of phase 1: applied by the view that holds the key's partition, then cleared
        const unsigned long long n = hpoly[4 + tid], key = hpoly[tid];
        Probe p;
        if (n && probe_home(t.g, key, p)) table_update<false>(t, key, (unsigned)n, claimed);
        hpoly[4 + tid] = 0ull;
    }
    __shared__ unsigned long long s_w;
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned long long* f_key = reinterpret_cast<unsigned long long*>(dyn);      // fold table of one chunk:
    unsigned int* f_cnt = reinterpret_cast<unsigned int*>(f_key + RP_FOLD);      // key -> occurrences, open addressing
    if (FOLD) for (int i = tid; i < RP_FOLD; i += RP_THREADS) { f_key[i] = 0ull; f_cnt[i] = 0u; }
    unsigned long long* counters = chunk_start + nperm + 1;
    unsigned last_lp = 0xFFFFFFFFu;
    // a CTA serves its own group first and then helps the following groups finish
    for (unsigned gi = 0; gi < G; gi++) {
        const unsigned g = (blockIdx.x + gi) % G;
        const unsigned q_end = (g + 1) * per * nsrc;
        const unsigned long long w_end = chunk_start[q_end];
        unsigned q = g * per * nsrc;
        while (true) {
            __syncthreads();                       // everyone has read the previous s_w
            if (tid == 0) s_w = atomicAdd(&counters[g], 1ull);
            __syncthreads();
            const unsigned long long w = s_w;
            if (w >= w_end) break;
            while (chunk_start[q + 1] <= w) q++;
            unsigned lp, src;
            replay_segment(q, nsrc, nlocal, G, lp, src);
            const unsigned seg = src * nlocal + lp;
            const unsigned n = min(cursor[seg], cap);
            const unsigned long long* base = keys + (unsigned long long)seg * cap;
            const unsigned i0 = (unsigned)(w - chunk_start[q]) * RP_CHUNK;

            if (prefetch && lp != last_lp) {
                // first chunk this CTA sees of bin lp: pull its slice of the group's NEXT bin's partitions into L2
                last_lp = lp;
                if (tid == 0 && lp + G < nlocal) {
                    const unsigned long long gb = (unsigned long long)bin0 + lp + G;                 // global bin
                    const unsigned long long p_lo = gb * t.g.nparts / nbins_global;
                    unsigned long long p_hi = ((gb + 1) * t.g.nparts - 1) / nbins_global;
                    if (p_lo >= t.g.part0 && p_lo < (unsigned long long)t.g.part0 + t.g.nlocal) {
                        if (p_hi >= (unsigned long long)t.g.part0 + t.g.nlocal) p_hi = (unsigned long long)t.g.part0 + t.g.nlocal - 1;
                        const unsigned long long r0 = (p_lo - t.g.part0) * t.g.subcap * sizeof(Slot);
                        const unsigned long long r1 = (p_hi + 1 - t.g.part0) * t.g.subcap * sizeof(Slot);
                        const unsigned ctas = gridDim.x / G ? gridDim.x / G : 1u;
                        unsigned long long slice = ((r1 - r0) / ctas + 127ull) & ~127ull;
                        unsigned long long a = r0 + slice * (blockIdx.x / G), e = a + slice;
                        if (e > r1) e = r1;
                        const char* tb = reinterpret_cast<const char*>(t.slots);
                        for (; a < e; a += 16384) prefetch_l2_bulk(tb + a, (unsigned)((e - a) < 16384ull ? (e - a) : 16384ull));
                    }
                }
            }

            // ---- the chunk's duplicates folded in shared memory first.  An expressed transcript's k-mers arrive hundreds
            // of times per bin (and N GPUs send N times as many to the one owner): folded here, a key costs ONE global
            // atomic per chunk instead of one per occurrence, and same-address atomics stop serialising in L2.
#pragma unroll
            for (int u = 0; FOLD && u < RP_PER_THREAD; u++) {
                const unsigned i = i0 + u * RP_THREADS + tid;
                const unsigned long long key = i < n ? __ldcs(base + i) : 0ull;
                if (key != 0ull) {
                    // bits 40..: every key of a bin shares the top bits of the LOW hash word (they are the partition)
                    unsigned h = (unsigned)(mix64(key) >> 40) & (RP_FOLD - 1);
                    int tries = 0;
                    for (; tries < RP_FOLD_PROBES; tries++) {
                        const unsigned long long old = atomicCAS(&f_key[h], 0ull, key);
                        if (old == 0ull || old == key) { atomicAdd(&f_cnt[h], 1u); break; }
                        h = (h + 1) & (RP_FOLD - 1);
                    }
                    if (tries == RP_FOLD_PROBES) table_update<false>(t, key, 1u, claimed);      // crowded corner: unfolded
                }
            }
            if (FOLD) __syncthreads();
#pragma unroll 1
            for (int gg = 0; gg < RP_PER_THREAD; gg += RP_GROUP) {
                unsigned long long key[RP_GROUP], cur[RP_GROUP], cur1[RP_GROUP];
                unsigned cnt[RP_GROUP];
                Probe pr[RP_GROUP];
#pragma unroll
                for (int u = 0; u < RP_GROUP; u++) {
                    if (FOLD) {
                        const unsigned sidx = (gg + u) * RP_THREADS + tid;
                        key[u] = f_key[sidx];
                        cnt[u] = f_cnt[sidx];
                        if (key[u] != 0ull) { f_key[sidx] = 0ull; f_cnt[sidx] = 0u; }     // clean for the next chunk
                    } else {
                        const unsigned i = i0 + (gg + u) * RP_THREADS + tid;
                        key[u] = i < n ? __ldcs(base + i) : 0ull;
                        cnt[u] = 1u;
                    }
                }
                // the first TWO slots of the home bucket in one 256-bit load: buckets fill front to back, so most keys
                // that are not in slot 0 are in slot 1 and need no second (dependent) round trip to L2
#pragma unroll
                for (int u = 0; u < RP_GROUP; u++)
                    if (key[u] != 0ull) {
                        if (probe_home(t.g, key[u], pr[u])) {
                            unsigned long long w0, w1;
                            ld_slot_pair(&t.slots[pr[u].base + pr[u].off], cur[u], w0, cur1[u], w1);
                        } else { key[u] = 0ull; atomicExch(t.error, 2); }
                    }
#pragma unroll
                for (int u = 0; u < RP_GROUP; u++)
                    if (key[u] != 0ull) {
                        if (cur[u] != key[u] && cur[u] != 0ull) {      // slot 0 holds another key: go on from slot 1
                            probe_next(t.g, pr[u]);
                            cur[u] = cur1[u];
                        }
                        Slot* sl = table_upsert_slot(t, key[u], pr[u], cur[u], claimed);
                        if (sl) atomicAdd(&sl->val, cnt[u]);
                    }
                __syncwarp();   // lanes leave the probe loops at different times: without this the warp stays split
                                // and every later log load / probe is issued once per lane subset
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
    if (lane == 0 && claimed) atomicAdd(t.n_claimed, (unsigned long long)claimed);
}

// ---- refine: coarse bins -> fine bins (owner side) ------------------------------------------------------------
// Phase 1 is only fast while a tile's entries fall into few bins (measured, 1.5 G entries: 256 bins 14.9 ms, 512 bins
// 17.7 ms, 1024 bins 36.6 ms, 4096 bins 63.8 ms -- a run of a bin inside a tile shrinks to one or two 8-B entries and
// every store becomes a partial-sector write).  So the exchange uses COARSE bins (a few hundred over all GPUs: long
// runs, efficient NVLink stores), and the owner splits each coarse bin into the f table partitions it covers with
// this kernel: a chunk of 2048 keys of one coarse segment is counting-sorted in shared memory by fine bin (only f
// bins occur, so runs are hundreds of entries) and streamed out; the source segments merge on the way, so the replay
// that follows sees one segment per fine bin.  Cost: 8 B read + 8 B written per entry, all sequential.
constexpr int RF_THREADS = 256;
constexpr int RF_PER_THREAD = RP_CHUNK / RF_THREADS;
static_assert(RP_CHUNK % RF_THREADS == 0, "chunk = whole rounds of the CTA");
constexpr unsigned RF_MAX_SPLIT = RF_THREADS;       // table partitions per coarse bin: one counter per thread

__global__ void __launch_bounds__(RF_THREADS, 4)
k_log_refine(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ cursor, unsigned cap,
             unsigned nsrc, unsigned ncoarse, const unsigned long long* __restrict__ chunk_start,
             unsigned long long* __restrict__ out_keys, unsigned int* __restrict__ out_cursor, unsigned out_cap,
             unsigned nfine, unsigned fine0, unsigned nfine_global, int* error, TableView t) {
    unsigned claimed = 0;
    // a chunk belongs to ONE coarse bin, so only its f = nfine / ncoarse fine bins can occur: all bookkeeping is per f
    __shared__ unsigned long long skey[RP_CHUNK];
    __shared__ unsigned short sbin[RP_CHUNK];
    __shared__ unsigned int cnt[RF_MAX_SPLIT], delta[RF_MAX_SPLIT], off[RF_MAX_SPLIT];
    __shared__ unsigned wtot[RF_THREADS / 32];
    __shared__ unsigned s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned f = nfine / ncoarse;
    const unsigned nseg = nsrc * ncoarse;
    const unsigned long long nchunks = chunk_start[nseg];
    cnt[tid] = 0u;
    __syncthreads();
    unsigned q = 0;
    for (unsigned long long w = blockIdx.x; w < nchunks; w += gridDim.x) {
        while (chunk_start[q + 1] <= w) q++;                       // segments in plan order: (coarse bin, source)
        const unsigned lb = q / nsrc, src = q % nsrc, seg = src * ncoarse + lb;
        const unsigned n = min(cursor[seg], cap);
        const unsigned long long* base = keys + (unsigned long long)seg * cap;
        const unsigned i0 = (unsigned)(w - chunk_start[q]) * RP_CHUNK;
        const unsigned first = fine0 + lb * f;                     // global index of this coarse bin's first partition
        // ---- A: fine bin (inside the coarse bin) and rank of every key
        unsigned long long key[RF_PER_THREAD];
        unsigned meta[RF_PER_THREAD];                              // bin << 12 | rank, ~0 = no entry
#pragma unroll
        for (int j = 0; j < RF_PER_THREAD; j++) {
            const unsigned i = i0 + j * RF_THREADS + tid;
            key[j] = i < n ? __ldcs(base + i) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < RF_PER_THREAD; j++) {
            meta[j] = 0xFFFFFFFFu;
            if (key[j] != 0ull) {
                const unsigned bin = hash_part(mix64(key[j]), nfine_global) - first;
                if (bin < f) meta[j] = (bin << 12) | atomicAdd(&cnt[bin], 1u);
                else atomicExch(error, 2);                         // a key that does not belong to this coarse bin
            }
        }
        __syncthreads();
        // ---- S: scan the f counters (one per thread), reserve every bin's run in the fine log
        {
            const unsigned c = (unsigned)tid < f ? cnt[tid] : 0u;
            unsigned incl = c;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) wtot[warp] = incl;
            __syncthreads();
            unsigned run = incl - c;
            for (int ww = 0; ww < warp; ww++) run += wtot[ww];
            if (tid == RF_THREADS - 1) s_total = run + c;
            off[tid] = run;
            if (c) delta[tid] = atomicAdd(&out_cursor[lb * f + tid], c) - run;
            cnt[tid] = 0u;
        }
        __syncthreads();
        // ---- B: keys into their sorted places
#pragma unroll
        for (int j = 0; j < RF_PER_THREAD; j++)
            if (meta[j] != 0xFFFFFFFFu) {
                const unsigned bin = meta[j] >> 12, idx = off[bin] + (meta[j] & 0xFFFu);
                skey[idx] = key[j];
                sbin[idx] = (unsigned short)bin;
            }
        __syncthreads();
        // ---- W: stream the sorted chunk out: runs of hundreds of entries per fine bin
        const unsigned total = s_total;
        for (unsigned i = tid; i < total; i += RF_THREADS) {
            const unsigned bin = sbin[i], pos = delta[bin] + i;
            if (pos < out_cap) out_keys[(unsigned long long)(lb * f + bin) * out_cap + pos] = skey[i];
            else if (t.slots) table_update<false>(t, skey[i], 1u, claimed);      // fine bin full (a repeat k-mer): count directly
            else atomicExch(error, 3);
        }
        __syncthreads();
    }
    if (t.slots) {
        for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
        if (lane == 0 && claimed) atomicAdd(t.n_claimed, (unsigned long long)claimed);
    }
}

size_t log_refine_plan_words(unsigned nsrc, unsigned ncoarse) { return (size_t)nsrc * ncoarse + 2; }
unsigned log_refine_max_split() { return RF_MAX_SPLIT; }

cudaError_t launch_log_refine(const unsigned long long* d_keys, const unsigned int* d_cursor, unsigned cap, unsigned nsrc,
                              unsigned ncoarse, unsigned long long* d_chunk_start, unsigned long long* d_out_keys,
                              unsigned int* d_out_cursor, unsigned out_cap, unsigned nfine, unsigned fine0,
                              unsigned nfine_global, int* d_error, TableView t, int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_log_refine", s);
    if (nsrc == 0 || ncoarse == 0 || nfine == 0) return cudaSuccess;
    if (nfine % ncoarse || nfine / ncoarse > RF_MAX_SPLIT) return cudaErrorInvalidValue;
    k_log_plan<<<1, 1024, 0, s>>>(d_cursor, cap, nsrc, ncoarse, 1, d_chunk_start);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int grid = max_resident_ctas((const void*)k_log_refine, RF_THREADS, 0, -1);
    if (grid <= 0) grid = sm_count;
    k_log_refine<<<grid, RF_THREADS, 0, s>>>(d_keys, d_cursor, cap, nsrc, ncoarse, d_chunk_start, d_out_keys, d_out_cursor,
                                             out_cap, nfine, fine0, nfine_global, d_error, t);
    return cudaGetLastError();
}

size_t log_replay_plan_words(unsigned nsrc, unsigned nlocal, unsigned G) {
    return (size_t)replay_per_group(nlocal, G) * G * nsrc + 1 + G;
}

cudaError_t launch_log_replay(const unsigned long long* d_keys, const unsigned int* d_cursor, unsigned cap, unsigned nsrc,
                              unsigned nlocal, unsigned bin0, unsigned nbins_global, unsigned groups,
                              unsigned long long* d_chunk_start, unsigned long long* d_hpoly, TableView t, int prefetch,
                              int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_log_replay", s);
    if (nsrc == 0 || nlocal == 0) return cudaSuccess;
    if (groups < 1) groups = 1;
    if (groups > nlocal) groups = nlocal;
    if (groups > 64) groups = 64;
    k_log_plan<<<1, 1024, 0, s>>>(d_cursor, cap, nsrc, nlocal, groups, d_chunk_start);
    k_log_plan_counters<<<1, 64, 0, s>>>(nsrc, nlocal, groups, d_chunk_start);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const bool fold = (prefetch & 2) != 0;             // bit 1 of the flags word: fold duplicates per chunk
    const size_t dyn = fold ? (size_t)RP_FOLD * 12 : 0;
    const void* kern = fold ? (const void*)k_log_replay<true> : (const void*)k_log_replay<false>;
    int grid = max_resident_ctas(kern, RP_THREADS, dyn, -1);
    if (grid <= 0) grid = sm_count;
    grid = grid / (int)groups * (int)groups;           // the same number of CTAs in every group
    if (grid < (int)groups) grid = (int)groups;
    if (fold)
        k_log_replay<true><<<grid, RP_THREADS, dyn, s>>>(d_keys, d_cursor, cap, nsrc, nlocal, bin0, nbins_global, groups,
                                                       d_chunk_start, d_hpoly, t, prefetch & 1);
    else
        k_log_replay<false><<<grid, RP_THREADS, dyn, s>>>(d_keys, d_cursor, cap, nsrc, nlocal, bin0, nbins_global, groups,
                                                        d_chunk_start, d_hpoly, t, prefetch & 1);
    return cudaGetLastError();
}

// =========================================================================================================
// (packed key, value) pairs and rehash
// =========================================================================================================
template <bool IS_MAX>
__global__ void __launch_bounds__(256)
k_load_pairs(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, int k,
             int canonical, TableView t) {
    unsigned claimed = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned p0, p1;
        packed_to_planes(keys[i], k, p0, p1);
        unsigned long long key = make_key(p0, p1);
        const unsigned long long kr = make_key(rc_plane(p0, k), rc_plane(p1, k));
        if (IS_MAX) {          // label table: (forward k-mer, bundle index + 1) -> the field of its orientation
            if (table_label_max(t, kr < key ? kr : key, kr < key, vals[i])) claimed++;
        } else {
            if (canonical) key = kr < key ? kr : key;
            table_update<false>(t, key, vals[i], claimed);
        }
    }
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(t.n_claimed, (unsigned long long)claimed);
}

cudaError_t launch_load_pairs(const uint64_t* d_keys, const uint32_t* d_vals, uint64_t n, int k, int canonical,
                              TableView t, int is_label, cudaStream_t s) {
    TimedLaunch timed("k_load_pairs", s);
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (is_label) k_load_pairs<true><<<(int)blocks, 256, 0, s>>>(d_keys, d_vals, n, k, canonical, t);
    else k_load_pairs<false><<<(int)blocks, 256, 0, s>>>(d_keys, d_vals, n, k, canonical, t);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
k_rehash(const Slot* __restrict__ from, uint64_t from_cap, TableView to, int is_label, uint32_t min_val) {
    unsigned claimed = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < from_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = __ldcs(reinterpret_cast<const uint4*>(&from[i]));
        const unsigned long long key = ((unsigned long long)s.y << 32) | s.x;
        if (key == 0ull) continue;
        if (is_label) {        // both orientations' labels move with the key
            if (s.z && table_label_max(to, key, false, s.z)) claimed++;
            if (s.w && table_label_max(to, key, true, s.w)) claimed++;
        } else if (s.z >= min_val) {
            table_update<false>(to, key, s.z, claimed);
        }
    }
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(to.n_claimed, (unsigned long long)claimed);
}

// direct-mapped cache fill: every live slot with val >= min_val goes to hot[mulhi(hash, hot_cap)] if that place is free
__global__ void __launch_bounds__(256)
k_hot_fill(const Slot* __restrict__ from, uint64_t from_cap, Slot* hot, uint64_t hot_cap, uint32_t min_val, uint32_t max_val) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < from_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = __ldcs(reinterpret_cast<const uint4*>(&from[i]));
        const unsigned long long key = ((unsigned long long)s.y << 32) | s.x;
        if (key == 0ull || s.z < min_val || s.z > max_val) continue;
        Slot* h = &hot[__umul64hi(mix64(key), hot_cap)];
        if (atomicCAS(&h->key, 0ull, key) == 0ull) h->val = s.z;
    }
}

cudaError_t launch_hot_fill(const Slot* from, uint64_t from_cap, Slot* hot, uint64_t hot_cap, uint32_t min_val,
                            uint32_t max_val, cudaStream_t s) {
    TimedLaunch timed("k_hot_fill", s);
    if (from_cap == 0 || hot_cap == 0) return cudaSuccess;
    uint64_t blocks = (from_cap + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_hot_fill<<<(int)blocks, 256, 0, s>>>(from, from_cap, hot, hot_cap, min_val, max_val);
    return cudaGetLastError();
}

cudaError_t launch_rehash(const Slot* from, uint64_t from_cap, TableView to, int is_label, uint32_t min_val,
                          cudaStream_t s) {
    TimedLaunch timed("k_rehash", s);
    if (from_cap == 0) return cudaSuccess;
    uint64_t blocks = (from_cap + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_rehash<<<(int)blocks, 256, 0, s>>>(from, from_cap, to, is_label, min_val);
    return cudaGetLastError();
}

// =========================================================================================================
// Per-read machinery shared by stats and assign.  GS = group size: 32 (one warp per read, buffers in shared
// memory) or LONG_THREADS (one CTA per read, buffers in global scratch).
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

// Median of n <= PR_MAXWIN u32 values held in shared memory, by one warp, WITHOUT sorting: a bisection on the VALUE
// between the warp minimum and maximum (coverage values of one read sit in a narrow band, so a handful of rounds),
// each round one compare per element and one redux.sync.  PER = elements per lane = ceil(n / 32), a template
// parameter so that a 100-bp read (76 windows, PER = 3) does not pay for the 256-window maximum.  Returns
// median_coverage() of fastaToKmerCoverageStats.cpp:337-347: odd n -> the middle element, even n -> the (wrapping)
// u32 mean of the two middle elements.
template <int PER>
__device__ __forceinline__ uint32_t warp_median_per(const uint32_t* __restrict__ v, int n, int lane) {
    unsigned x[PER];
    const bool tail_live = (PER - 1) * 32 + lane < n;      // only the last element of a lane can lie past n
    unsigned mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const bool live = i < PER - 1 || tail_live;
        x[i] = live ? v[i * 32 + lane] : 0u;
        if (live) { mn = min(mn, x[i]); mx = max(mx, x[i]); }
    }
    unsigned lo = __reduce_min_sync(FULL, mn), hi = __reduce_max_sync(FULL, mx);
    const unsigned k1 = (unsigned)(n - 1) / 2u, k2 = (unsigned)n / 2u;
    // smallest value with at least k1 + 1 elements <= it = the element of rank k1
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        unsigned cnt = 0;
#pragma unroll
        for (int i = 0; i < PER; i++) cnt += ((i < PER - 1 || tail_live) && x[i] <= mid) ? 1u : 0u;
        cnt = __reduce_add_sync(FULL, cnt);
        if (cnt >= k1 + 1u) hi = mid; else lo = mid + 1u;
    }
    const unsigned x1 = lo;
    if (k1 == k2) return x1;
    unsigned le = 0, nxt = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const bool live = i < PER - 1 || tail_live;
        le += (live && x[i] <= x1) ? 1u : 0u;
        if (live && x[i] > x1) nxt = min(nxt, x[i]);
    }
    le = __reduce_add_sync(FULL, le);
    nxt = __reduce_min_sync(FULL, nxt);
    const unsigned x2 = le >= k2 + 1u ? x1 : nxt;          // the element of rank k2 = k1 + 1
    return (uint32_t)(x1 + x2) / 2u;
}

__device__ __forceinline__ uint32_t warp_median_u32(const uint32_t* __restrict__ v, int n, int lane) {
    static_assert(PR_MAXWIN == 256, "dispatch below covers 1..8 elements per lane");
    switch ((n + 31) >> 5) {           // exact: warp_median_per assumes only a lane's LAST element can lie past n
        case 0: case 1: return warp_median_per<1>(v, n, lane);
        case 2: return warp_median_per<2>(v, n, lane);
        case 3: return warp_median_per<3>(v, n, lane);
        case 4: return warp_median_per<4>(v, n, lane);
        case 5: return warp_median_per<5>(v, n, lane);
        case 6: return warp_median_per<6>(v, n, lane);
        case 7: return warp_median_per<7>(v, n, lane);
        default: return warp_median_per<8>(v, n, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------
// coverage statistics of one read (fastaToKmerCoverageStats.cpp:300-402)
// ---------------------------------------------------------------------------------------------------------
template <int GS>
__device__ __forceinline__ void read_cov_stats(const uint8_t* __restrict__ seq, int L, int k, int canonical,
                                               const Slot* __restrict__ slots, Geo geo, uint32_t* P0,
                                               uint32_t* P1, uint32_t* PB, uint32_t* cov, float* sq,
                                               unsigned long long* red, uint32_t* per_kmer, uint32_t& median,
                                               float& mean, float& stdev, int gtid) {
    const int nwin = L >= k ? L - k + 1 : 0;
    if (nwin == 0) {   // S6: shorter than k -> empty vector; S7-S9 on n = 0: 0, 0, sqrt(0/-1) = -0
        median = 0; mean = 0.0f; stdev = __int_as_float(0x80000000);
        return;
    }
    const unsigned mk = kmask(k);
    const int nch = (L + 31) >> 5;
    pack_read_planes<GS>(seq, L, nch, P0, P1, PB, gtid);
    gsync<GS>();

    unsigned long long part = 0;
    for (int pb = 0; pb < nwin; pb += GS) {     // every lane runs every iteration: table_lookup is warp-convergent
        const int p = pb + gtid;
        const bool live = p < nwin;
        bool ok = false;
        unsigned long long key = 0ull;
        if (live) {
            const int c = p >> 5, o = p & 31;
            const unsigned bad = __funnelshift_r(PB[c], PB[c + 1], o) & mk;
            if (!bad) {
                const unsigned f0 = __funnelshift_r(P0[c], P0[c + 1], o) & mk;
                const unsigned f1 = __funnelshift_r(P1[c], P1[c + 1], o) & mk;
                key = make_key(f0, f1);
                if (canonical) {
                    const unsigned long long kr = make_key(rc_plane(f0, k), rc_plane(f1, k));
                    key = kr < key ? kr : key;
                }
                ok = true;
            }
        }
        unsigned v = table_lookup(slots, geo, key, ok);
        if (live) {
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
            // strict read order; four squares per load where sq is 16-B aligned (always on the warp path)
            const float4* sq4 = reinterpret_cast<const float4*>(sq);
            int p = 0;
            for (; (reinterpret_cast<unsigned long long>(sq) & 15ull) == 0ull && p + 4 <= nwin; p += 4) {
                const float4 q = sq4[p >> 2];
                acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, q.x), q.y), q.z), q.w);
            }
            for (; p < nwin; p++) acc = __fadd_rn(acc, sq[p]);
            acc = __fsqrt_rn(__fdiv_rn(acc, __int2float_rn(nwin - 1)));
        }
        sd = acc;
    }
    // median: odd -> middle, even -> u32 (wrapping) mean of the two middles
    if (GS == 32) {
        median = warp_median_u32(cov, nwin, gtid);         // selection, no sort
    } else {
        const unsigned n2 = next_pow2((unsigned)nwin);
        for (unsigned p = nwin + gtid; p < n2; p += GS) cov[p] = 0xFFFFFFFFu;
        gsync<GS>();
        bitonic_sort<GS, uint32_t>(cov, n2, gtid);
        median = (nwin & 1) ? cov[nwin / 2] : (uint32_t)(cov[(nwin - 1) / 2] + cov[nwin / 2]) / 2u;
    }
    mean = avg;
    stdev = sd;    // meaningful in gtid 0 only
}

struct alignas(16) PerReadSmem {
    uint32_t p0[PR_WARPS][PR_MAXCH];
    uint32_t p1[PR_WARPS][PR_MAXCH];
    uint32_t pb[PR_WARPS][PR_MAXCH];
    uint32_t a[PR_WARPS][2 * PR_MAXWIN];    // stats: cov[PR_MAXWIN] + sq[PR_MAXWIN]; assign: hits[2*PR_MAXWIN]
    unsigned int nhits[PR_WARPS];
};

__global__ void __launch_bounds__(PR_WARPS * 32, 6)
k_cov_stats(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, uint64_t nreads,
            int k, int canonical, const Slot* __restrict__ slots, Geo geo, uint32_t* __restrict__ median,
            float* __restrict__ mean, float* __restrict__ stdev, uint32_t* __restrict__ per_kmer, LongList ll) {
    __shared__ PerReadSmem sm;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t r = (uint64_t)blockIdx.x * PR_WARPS + w;
    if (r >= nreads) return;
    const uint64_t o0 = offs[r], o1 = offs[r + 1];
    const int L = (int)(o1 - o0 - 1);            // the record's last byte is its '\n' terminator
    const int nwin = L >= k ? L - k + 1 : 0;
    if (nwin > PR_MAXWIN) {
        if (lane == 0) {
            const unsigned slot = atomicAdd(ll.count, 1u);
            ll.idx[slot] = (unsigned)r;
            atomicMax(ll.max_win, (unsigned)nwin);
        }
        return;
    }
    uint32_t med; float mu, sd;
    read_cov_stats<32>(recs + (o0 - rec_base), L, k, canonical, slots, geo, sm.p0[w], sm.p1[w], sm.pb[w], sm.a[w],
                       reinterpret_cast<float*>(sm.a[w] + PR_MAXWIN), nullptr,
                       per_kmer ? per_kmer + (o0 - rec_base) : nullptr, med, mu, sd, lane);
    if (lane == 0) { median[r] = med; mean[r] = mu; stdev[r] = sd; }
}

cudaError_t launch_cov_stats(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                             int canonical, const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean,
                             float* d_stdev, uint32_t* d_per_kmer, LongList ll, cudaStream_t s) {
    TimedLaunch timed("k_cov_stats", s);
    if (nreads == 0) return cudaSuccess;
    const uint64_t blocks = (nreads + PR_WARPS - 1) / PR_WARPS;
    k_cov_stats<<<(unsigned)blocks, PR_WARPS * 32, 0, s>>>(d_recs, d_offs, rec_base, nreads, k, canonical, slots, geo,
                                                           d_median, d_mean, d_stdev, d_per_kmer, ll);
    return cudaGetLastError();
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
                 unsigned long long scratch_words, int* error) {
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
        const uint64_t r = long_idx[i];
        const uint64_t o0 = offs[r], o1 = offs[r + 1];
        const int L = (int)(o1 - o0 - 1);
        uint32_t med; float mu, sd;
        read_cov_stats<LONG_THREADS>(recs + (o0 - rec_base), L, k, canonical, slots, geo, P0, P1, PB, cov, sq, red,
                                     per_kmer ? per_kmer + (o0 - rec_base) : nullptr, med, mu, sd, threadIdx.x);
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
                                                    nullptr);
    return cudaGetLastError();
}

cudaError_t launch_cov_stats_long_auto(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k,
                                       int canonical, const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean,
                                       float* d_stdev, uint32_t* d_per_kmer, LongList ll, void* d_scratch,
                                       size_t scratch_bytes, int* d_error, int nctas, cudaStream_t s) {
    TimedLaunch timed("k_cov_stats_long", s);
    k_cov_stats_long<<<nctas, LONG_THREADS, 0, s>>>(d_recs, d_offs, rec_base, k, canonical, slots, geo, d_median, d_mean,
                                                    d_stdev, d_per_kmer, ll.idx, 0, 0, (uint32_t*)d_scratch, 0, ll.count,
                                                    scratch_bytes / 4, d_error);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// read -> bundle vote of one read (ReadsToTranscripts.cc:216-274)
// ---------------------------------------------------------------------------------------------------------
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
    unsigned nh = 0;                                // GS == 32: hits appended so far (warp-uniform)
    for (int pb = 0; pb < nwin; pb += GS) {         // every lane runs every iteration: table_lookup is warp-convergent
        const int p = pb + gtid;
        bool do_f = false, do_r = false, is_rc = false, pal = false;
        unsigned long long key = 0ull;
        if (p < nwin) {
            const int c = p >> 5, o = p & 31;
            const unsigned bad = __funnelshift_r(PB[c], PB[c + 1], o) & mk;
            if (!bad) {                     // a window with a non-ACGT character can never equal a table k-mer
                const unsigned f0 = __funnelshift_r(P0[c], P0[c + 1], o) & mk;
                const unsigned f1 = __funnelshift_r(P1[c], P1[c + 1], o) & mk;
                do_f = window_entropy_ok(lut, f0, f1, mk, false);
                do_r = !strand && window_entropy_ok(lut, f0, f1, mk, true);
                const unsigned long long kf = make_key(f0, f1), kr = make_key(rc_plane(f0, k), rc_plane(f1, k));
                is_rc = kr < kf; pal = kr == kf;
                key = is_rc ? kr : kf;
            }
        }
        // ONE probe answers both passes of the reference (forward window, then reverse-complemented window): the slot
        // of the canonical key holds the label of the bundle k-mer equal to the key (val) and of the bundle k-mer whose
        // reverse complement is the key (aux).  A palindrome (even k only) is its own reverse complement.
        const uint2 v = table_lookup2(slots, geo, key, do_f || do_r);
        const unsigned vf = do_f ? (is_rc ? v.y : v.x) : 0u;
        const unsigned vr = do_r ? ((is_rc || pal) ? v.x : v.y) : 0u;
        if (GS == 32) {
            const unsigned lt = (1u << gtid) - 1u;
            const unsigned mf = __ballot_sync(FULL, vf != 0u), mr = __ballot_sync(FULL, vr != 0u);
            if (vf) hits[nh + __popc(mf & lt)] = (int32_t)vf - 1;
            nh += __popc(mf);
            if (vr) hits[nh + __popc(mr & lt)] = (int32_t)vr - 1;
            nh += __popc(mr);
        } else {
            if (vf) hits[atomicAdd(nhits_p, 1u)] = (int32_t)vf - 1;
            if (vr) hits[atomicAdd(nhits_p, 1u)] = (int32_t)vr - 1;
        }
    }
    gsync<GS>();
    const int n = GS == 32 ? (int)nh : (int)*nhits_p;
    int b = -1, sc = 0;
    if (n >= 2 && GS == 32) {
        // The reference sorts the hits and scans the runs (ReadsToTranscripts.cc:253-268): a label with m hits scores
        // m-1, the largest label m-2, strict '>' while ascending => ties go to the smaller label.  The same result
        // without a sort: walk the DISTINCT labels in ascending order (almost always one or two), one warp min and
        // one warp count per label.
        int last = -1;
        for (int p = gtid; p < n; p += 32) last = max(last, hits[p]);
        last = __reduce_max_sync(FULL, last);
        int cur = -1;
        while (true) {
            int mn = 0x7FFFFFFF;
            for (int p = gtid; p < n; p += 32) { const int h = hits[p]; if (h > cur) mn = min(mn, h); }
            const int L = __reduce_min_sync(FULL, mn);
            if (L == 0x7FFFFFFF) break;
            unsigned m = 0;
            for (int p = gtid; p < n; p += 32) m += hits[p] == L ? 1u : 0u;
            m = __reduce_add_sync(FULL, m);
            const int s = (int)m - 1 - (L == last ? 1 : 0);
            if (s > sc) { sc = s; b = L; }
            cur = L;
        }
        if (sc <= 0) { b = -1; sc = 0; }
    } else if (n >= 2) {
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
        if (GS > 32) {
            __shared__ int wsc[LONG_THREADS / 32], wb[LONG_THREADS / 32];
            gsync<GS>();
            if ((gtid & 31) == 0) { wsc[gtid >> 5] = sc; wb[gtid >> 5] = b; }
            gsync<GS>();
            sc = wsc[0]; b = wb[0];
            for (int w = 1; w < GS / 32; w++)
                if (wsc[w] > sc || (wsc[w] == sc && wsc[w] > 0 && wb[w] < b)) { sc = wsc[w]; b = wb[w]; }
            gsync<GS>();
        }
        if (sc <= 0) { b = -1; sc = 0; }
    }
    best = b; score = sc;
    // pct = (int)((float)max / num_kmer_pos * 100 + 0.5): fp32 divide, fp32 multiply, double add, truncate
    const float q = __fmul_rn(__fdiv_rn(__int2float_rn(sc), __int2float_rn(nwin)), 100.0f);
    pct = (int)__dadd_rn((double)q, 0.5);
}

__global__ void __launch_bounds__(PR_WARPS * 32)
k_assign(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, uint64_t rec_base, uint64_t nreads, int k,
         int strand, const Slot* __restrict__ slots, Geo geo, const uint8_t* __restrict__ lut,
         int32_t* __restrict__ best, int32_t* __restrict__ pct, int32_t* __restrict__ score, LongList ll) {
    __shared__ PerReadSmem sm;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t r = (uint64_t)blockIdx.x * PR_WARPS + w;
    if (r >= nreads) return;
    const uint64_t o0 = offs[r], o1 = offs[r + 1];
    const int L = (int)(o1 - o0 - 1);
    const int nwin = L - k + 1;
    if (nwin > PR_MAXWIN) {
        if (lane == 0) {
            const unsigned slot = atomicAdd(ll.count, 1u);
            ll.idx[slot] = (unsigned)r;
            atomicMax(ll.max_win, (unsigned)nwin);
        }
        return;
    }
    int32_t b, sc, pc;
    read_assign<32>(recs + (o0 - rec_base), L, k, strand, slots, geo, lut, sm.p0[w], sm.p1[w], sm.pb[w],
                    reinterpret_cast<int32_t*>(sm.a[w]), &sm.nhits[w], b, sc, pc, lane);
    if (lane == 0) { best[r] = b; pct[r] = pc; if (score) score[r] = sc; }
}

cudaError_t launch_assign(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                          int strand, const Slot* slots, Geo geo, const uint8_t* d_entropy_ok, int32_t* d_best,
                          int32_t* d_pct, int32_t* d_score, LongList ll, cudaStream_t s) {
    TimedLaunch timed("k_assign", s);
    if (nreads == 0) return cudaSuccess;
    const uint64_t blocks = (nreads + PR_WARPS - 1) / PR_WARPS;
    k_assign<<<(unsigned)blocks, PR_WARPS * 32, 0, s>>>(d_recs, d_offs, rec_base, nreads, k, strand, slots, geo,
                                                        d_entropy_ok, d_best, d_pct, d_score, ll);
    return cudaGetLastError();
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
        const uint64_t r = long_idx[i];
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
// Table scans: histo (jellyfish histo: bins 1..10000, 10001 = everything larger) and export (dump)
// =========================================================================================================
constexpr int HISTO_BINS = 10002;

__global__ void __launch_bounds__(256)
k_histo(const Slot* __restrict__ slots, uint64_t cap, unsigned long long* __restrict__ bins) {
    __shared__ unsigned int sb[HISTO_BINS];
    for (int i = threadIdx.x; i < HISTO_BINS; i += blockDim.x) sb[i] = 0;
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = __ldcs(reinterpret_cast<const uint4*>(&slots[i]));
        if ((s.x | s.y) == 0u) continue;
        const unsigned c = s.z;
        atomicAdd(&sb[c > 10000u ? 10001u : c], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HISTO_BINS; i += blockDim.x)
        if (sb[i]) atomicAdd(&bins[i], (unsigned long long)sb[i]);
}

cudaError_t launch_histo(const Slot* slots, uint64_t cap, unsigned long long* d_bins, cudaStream_t s) {
    TimedLaunch timed("k_histo", s);
    if (cap == 0) return cudaSuccess;
    uint64_t blocks = (cap + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    k_histo<<<(int)blocks, 256, 0, s>>>(slots, cap, d_bins);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
k_export(const Slot* __restrict__ slots, uint64_t cap, uint32_t min_count, uint32_t max_count, int k, int canonical_repr,
         uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals, unsigned long long* __restrict__ out_n) {
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (cap + stride - 1) / stride;
    for (uint64_t rd = 0; rd < rounds; rd++) {
        const uint64_t i = rd * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        bool keep = false;
        uint4 s = make_uint4(0, 0, 0, 0);
        if (i < cap) {
            s = __ldcs(reinterpret_cast<const uint4*>(&slots[i]));
            keep = (s.x | s.y) != 0u && s.z >= min_count && s.z <= max_count;
        }
        const unsigned m = __ballot_sync(FULL, keep);
        if (m == 0) continue;
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(out_n, (unsigned long long)__popc(m));
        basei = __shfl_sync(FULL, basei, 0);
        if (keep) {
            const uint64_t o = basei + __popc(m & ((1u << lane) - 1u));
            if (out_keys) {
                unsigned long long pk = planes_to_packed(s.x, s.y & 0x7FFFFFFFu, k);
                if (canonical_repr) {           // jellyfish prints the lexicographically smaller strand (A<C<G<T)
                    const unsigned long long rc = packed_revcomp(pk, k);
                    pk = rc < pk ? rc : pk;
                }
                out_keys[o] = pk;
                out_vals[o] = s.z;
            }
        }
    }
}

cudaError_t launch_export(const Slot* slots, uint64_t cap, uint32_t min_count, uint32_t max_count, int k,
                          int canonical_repr, uint64_t* d_keys, uint32_t* d_vals, unsigned long long* d_n,
                          cudaStream_t s) {
    TimedLaunch timed("k_export", s);
    if (cap == 0) return cudaSuccess;
    uint64_t blocks = (cap + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_export<<<(int)blocks, 256, 0, s>>>(slots, cap, min_count, max_count, k, canonical_repr, d_keys, d_vals, d_n);
    return cudaGetLastError();
}

// =========================================================================================================
// GUPS: the measured random-access roofline for this table geometry
// =========================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256)
k_gups(Slot* slots, uint64_t cap, uint64_t nops, unsigned long long* sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (uint64_t i = tid; i < nops; i += 4 * stride) {
        unsigned long long idx[4];
#pragma unroll
        for (int u = 0; u < 4; u++) idx[u] = __umul64hi(mix64(0x9E3779B97F4A7C15ull * (i + u * stride + 1)), cap);
        if (MODE == 0) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i + u * stride < nops) v[u] = __ldcg(reinterpret_cast<const uint4*>(&slots[idx[u]]));
                else v[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; u++) acc += v[u].x + v[u].z;
        } else {
            unsigned long long kv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                kv[u] = 0;
                if (i + u * stride < nops) {
                    if (MODE == 1) kv[u] = __ldcg(&slots[idx[u]].key);
                    else kv[u] = atomicCAS(&slots[idx[u]].key, 0ull, KEY_TAG | idx[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i + u * stride < nops) { atomicAdd(&slots[idx[u]].val, 1u); acc += kv[u]; }
        }
    }
    if (acc == 0x123456789ull) *sink = acc;   // keep the loads alive
}

cudaError_t launch_gups(Slot* slots, uint64_t cap, uint64_t nops, int mode, unsigned long long* d_sink, int sm_count,
                        cudaStream_t s) {
    TimedLaunch timed("k_gups", s);
    const int grid = sm_count * 8;
    if (mode == 0) k_gups<0><<<grid, 256, 0, s>>>(slots, cap, nops, d_sink);
    else if (mode == 1) k_gups<1><<<grid, 256, 0, s>>>(slots, cap, nops, d_sink);
    else k_gups<2><<<grid, 256, 0, s>>>(slots, cap, nops, d_sink);
    return cudaGetLastError();
}

}  // namespace tg
