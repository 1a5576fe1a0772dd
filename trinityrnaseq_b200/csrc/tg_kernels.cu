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
                    unsigned long long kf = make_key_k(f0, f1, k);
                    if (MODE == MODE_COUNT) {
                        if (canonical) {
                            const unsigned long long kr = make_key_k(rc_plane(f0, k), rc_plane(f1, k), k);
                            kf = kr < kf ? kr : kf;
                        }
                    } else {
                        // a valid window lies inside one record, so label < nrec whenever we advance
                        if (!bad) while (g0 + s >= next_off) { label++; next_off = offs[label + 1]; }
                        lab[u] = first_index + label + 1;
                        // label tables are keyed canonically with one label per orientation (tg_device.cuh)
                        const unsigned long long kr = make_key_k(rc_plane(f0, k), rc_plane(f1, k), k);
                        if (kr < kf) { kf = kr; rcs |= 1u << u; }
                    }
                    key[u] = bad ? 0ull : kf;
                    cnt[u] = 1;
                    if (k == 32 && !bad && kf == 0ull) {       // poly-A 32-mer: the key that reads as "no key" (tg_device.cuh)
                        Slot* z = zero_key_slot(t.slots, t.g);
                        if (MODE == MODE_COUNT) { if (atomicAdd(&z->val, 1u) == 0u) claimed++; }
                        else if (atomicMax((rcs >> u) & 1u ? &z->aux : &z->val, lab[u]) == 0u) claimed++;
                    }
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

// QUERY = the same partitioning for ROUTED LOOKUPS (multi-GPU statistics against a sharded table that is not replicated):
// every valid window's key travels to its owner like a counted k-mer does -- homopolymers included, a lookup has no side
// channel -- and the position of the window in the record buffer is kept locally in lg.posidx at the same (bin, pos): the
// owner's answers come back in the log's layout and are scattered to those positions.
template <bool QUERY>
__global__ void __launch_bounds__(LT_THREADS, 2)
k_log_tiles(const uint8_t* __restrict__ recs, uint64_t ntiles, int k, int canonical, LogView lg, TableView t) {
    __shared__ LogSmem sm;
    extern __shared__ __align__(16) unsigned char dyn[];
    // dynamic: skey[LT_TILE] u64 | delta[nbins] u32 | sbin[LT_TILE] u16 | cnt16[nbins2] u16 | off16[nbins2] u16 | QUERY: spos[LT_TILE] u16
    const unsigned nbins = lg.nbins, nbins2 = (nbins + 1u) & ~1u;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(dyn);
    unsigned int* delta = reinterpret_cast<unsigned int*>(skey + LT_TILE);
    unsigned short* sbin = reinterpret_cast<unsigned short*>(delta + nbins);
    unsigned short* cnt16 = sbin + LT_TILE;
    unsigned short* off16 = cnt16 + nbins2;
    unsigned short* spos = off16 + nbins2;
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
            const bool homo = !QUERY && (f0 == 0u || f0 == mk) && (f1 == 0u || f1 == mk);
            unsigned m = 0xFFFFFFFFu;
            if (!bad) {
                if (homo) {
                    const unsigned code = (f0 & 1u) | ((f1 & 1u) << 1);
                    hpA += code == 0u; hpC += code == 1u; hpG += code == 2u; hpT += code == 3u;
                } else {
                    const unsigned bin = hash_part(key_hash(window_key(f0, f1, k, canonical)), nbins);
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
                if (QUERY) spos[idx] = (unsigned short)(tid * LT_WIN + j);
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
                if (QUERY) lg.posidx[(unsigned long long)bin * lg.cap + pos] = (unsigned)(tile * LT_TILE + spos[i]);
            } else if (!QUERY && t.slots) {      // bin full: count this occurrence directly
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

size_t log_tiles_smem_bytes(unsigned nbins, bool query) {
    const unsigned nbins2 = (nbins + 1u) & ~1u;
    return (size_t)LT_TILE * 8 + (size_t)nbins * 4 + (size_t)LT_TILE * 2 + (size_t)nbins2 * 2 * 2 + (query ? (size_t)LT_TILE * 2 : 0);
}

cudaError_t launch_log_tiles(const uint8_t* d_recs, uint64_t nbytes, int k, int canonical, LogView lg, TableView t,
                             int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_log_tiles", s);
    if (nbytes == 0) return cudaSuccess;
    if (lg.nbins > LOG_MAX_BINS) return cudaErrorInvalidValue;
    if (reinterpret_cast<uintptr_t>(d_recs) & 15u) return cudaErrorMisalignedAddress;      // TMA bulk copies want 16-byte aligned tiles
    const uint64_t ntiles = (nbytes + LT_TILE - 1) / LT_TILE;
    const bool query = lg.posidx != nullptr;
    if (query && nbytes > 0xFFFF0000ull) return cudaErrorInvalidValue;      // return addresses are 32-bit buffer positions
    const size_t dyn = log_tiles_smem_bytes(lg.nbins, query);
    const void* kern = query ? (const void*)k_log_tiles<true> : (const void*)k_log_tiles<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LT_THREADS, dyn);
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)per_sm * sm_count;
    if (grid > ntiles) grid = ntiles;
    if (query) k_log_tiles<true><<<(unsigned)grid, LT_THREADS, dyn, s>>>(d_recs, ntiles, k, canonical, lg, t);
    else k_log_tiles<false><<<(unsigned)grid, LT_THREADS, dyn, s>>>(d_recs, ntiles, k, canonical, lg, t);
    return cudaGetLastError();
}

// ---- routed lookups, owner side and way back -------------------------------------------------------------------
// k_query_answer: the owner's receive log [nsrc][lp][cap] of QUERY keys -> resp[same layout] = the table's value (0 = absent).
// Segments are visited bin-major (all sources of a bin together), so that the CTAs in flight probe the partitions of a few
// neighbouring bins and those stay in L2 -- the replay's trick, without its counters.
__global__ void __launch_bounds__(256)
k_query_answer(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ cursor, unsigned cap, unsigned nsrc,
               unsigned lp, const Slot* __restrict__ slots, Geo geo, unsigned int* __restrict__ resp) {
    constexpr unsigned CHUNK = 2048;
    const unsigned chunks_per_seg = (cap + CHUNK - 1) / CHUNK;
    const unsigned long long nwork = (unsigned long long)nsrc * lp * chunks_per_seg;
    for (unsigned long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        const unsigned chunk = (unsigned)(w % chunks_per_seg);
        const unsigned long long sb = w / chunks_per_seg;              // bin-major: sb = bin * nsrc + src
        const unsigned src = (unsigned)(sb % nsrc), bin = (unsigned)(sb / nsrc);
        const unsigned seg = src * lp + bin;
        const unsigned n = min(cursor[seg], cap);
        const unsigned i0 = chunk * CHUNK;
        if (i0 >= n) continue;
        const unsigned long long base = (unsigned long long)seg * cap;
        for (unsigned i = i0 + threadIdx.x; i < min(i0 + CHUNK, (n + 31u) & ~31u); i += 256) {       // whole warps: the lookup is convergent
            const bool live = i < n;
            const unsigned long long key = live ? __ldcs(&keys[base + i]) : 0ull;
            const unsigned v = table_lookup(slots, geo, key, live && key != 0ull);
            if (live) resp[base + i] = v;
        }
    }
}
// k_query_scatter: answers back at the requester, in ITS log layout [nbins][cap] -> cov[position of the window]
__global__ void __launch_bounds__(256)
k_query_scatter(const unsigned int* __restrict__ resp, const unsigned int* __restrict__ posidx, const unsigned int* __restrict__ cursor,
                unsigned nbins, unsigned cap, unsigned int* __restrict__ cov) {
    constexpr unsigned CHUNK = 4096;
    const unsigned chunks_per_bin = (cap + CHUNK - 1) / CHUNK;
    const unsigned long long nwork = (unsigned long long)nbins * chunks_per_bin;
    for (unsigned long long w = blockIdx.x; w < nwork; w += gridDim.x) {
        const unsigned bin = (unsigned)(w / chunks_per_bin), i0 = (unsigned)(w % chunks_per_bin) * CHUNK;
        const unsigned n = min(cursor[bin], cap);
        const unsigned long long base = (unsigned long long)bin * cap;
        for (unsigned i = i0 + threadIdx.x; i < min(i0 + CHUNK, n); i += 256) cov[posidx[base + i]] = resp[base + i];
    }
}

cudaError_t launch_query_answer(const unsigned long long* d_keys, const unsigned int* d_cursor, unsigned cap, unsigned nsrc,
                                unsigned lp, const Slot* slots, Geo geo, unsigned int* d_resp, int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_query_answer", s);
    if (nsrc == 0 || lp == 0 || cap == 0) return cudaSuccess;
    k_query_answer<<<sm_count * 8, 256, 0, s>>>(d_keys, d_cursor, cap, nsrc, lp, slots, geo, d_resp);
    return cudaGetLastError();
}
cudaError_t launch_query_scatter(const unsigned int* d_resp, const unsigned int* d_posidx, const unsigned int* d_cursor,
                                 unsigned nbins, unsigned cap, unsigned int* d_cov, int sm_count, cudaStream_t s) {
    TimedLaunch timed("k_query_scatter", s);
    if (nbins == 0 || cap == 0) return cudaSuccess;
    k_query_scatter<<<sm_count * 8, 256, 0, s>>>(d_resp, d_posidx, d_cursor, nbins, cap, d_cov);
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
                    unsigned h = key_hash(key).y & (RP_FOLD - 1);
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
                const unsigned bin = hash_part(key_hash(key[j]), nfine_global) - first;
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
        unsigned long long key = make_key_k(p0, p1, k);
        const unsigned long long kr = make_key_k(rc_plane(p0, k), rc_plane(p1, k), k);
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

// Table scans cover slots [0, cap] -- cap is the zero-key slot (poly-A at k = 32, tg_device.cuh), which is live when it holds
// a value although its key reads as empty; at k <= 31 it stays all zero.
__device__ __forceinline__ bool slot_live(const uint4& s, uint64_t i, uint64_t cap) {
    return (s.x | s.y) != 0u || (i == cap && (s.z | s.w) != 0u);
}

// Only a fraction of the slots pass the filter of a compaction pass (and a third are occupied at all): the survivors of a
// warp's 32 slots are queued in shared memory and re-inserted 32 at a time, so that the probe + CAS + RED sequence always
// runs on full warps instead of on the few lanes whose slot qualified.
__global__ void __launch_bounds__(256)
k_rehash(const Slot* __restrict__ from, uint64_t from_cap, TableView to, int is_label, uint32_t min_val, uint32_t max_val) {
    __shared__ uint4 queue[8][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned claimed = 0, nq = 0;
    auto insert = [&](const uint4 s) {
        const unsigned long long key = ((unsigned long long)s.y << 32) | s.x;
        if (is_label) {        // both orientations' labels move with the key
            if (s.z && table_label_max(to, key, false, s.z)) claimed++;
            if (s.w && table_label_max(to, key, true, s.w)) claimed++;
        } else {
            table_update<false>(to, key, s.z, claimed);
        }
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (from_cap + 1 + stride - 1) / stride;
    for (uint64_t rd = 0; rd < rounds; rd++) {
        const uint64_t i = rd * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        uint4 s = make_uint4(0u, 0u, 0u, 0u);
        if (i <= from_cap) s = __ldcs(reinterpret_cast<const uint4*>(&from[i]));
        const bool keep = slot_live(s, i, from_cap) && (is_label || (s.z >= min_val && s.z <= max_val));
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) queue[w][nq + __popc(m & lt)] = s;
        nq += __popc(m);
        __syncwarp();
        if (nq >= 32u) {
            insert(queue[w][lane]);
            __syncwarp();
            nq -= 32u;
            const bool mv = (unsigned)lane < nq;               // the leftovers move to the front
            uint4 t2 = make_uint4(0u, 0u, 0u, 0u);
            if (mv) t2 = queue[w][32 + lane];
            __syncwarp();
            if (mv) queue[w][lane] = t2;
            __syncwarp();
        }
    }
    if ((unsigned)lane < nq) insert(queue[w][lane]);
    for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(FULL, claimed, o);
    if (lane == 0 && claimed) atomicAdd(to.n_claimed, (unsigned long long)claimed);
}

cudaError_t launch_rehash(const Slot* from, uint64_t from_cap, TableView to, int is_label, uint32_t min_val,
                          uint32_t max_val, cudaStream_t s) {
    TimedLaunch timed("k_rehash", s);
    if (from_cap == 0) return cudaSuccess;
    uint64_t blocks = (from_cap + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_rehash<<<(int)blocks, 256, 0, s>>>(from, from_cap, to, is_label, min_val, max_val);
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
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = __ldcs(reinterpret_cast<const uint4*>(&slots[i]));
        if (!slot_live(s, i, cap)) continue;
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
    const uint64_t rounds = (cap + 1 + stride - 1) / stride;
    for (uint64_t rd = 0; rd < rounds; rd++) {
        const uint64_t i = rd * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        bool keep = false;
        uint4 s = make_uint4(0, 0, 0, 0);
        if (i <= cap) {
            s = __ldcs(reinterpret_cast<const uint4*>(&slots[i]));
            keep = slot_live(s, i, cap) && s.z >= min_count && s.z <= max_count;
        }
        const unsigned m = __ballot_sync(FULL, keep);
        if (m == 0) continue;
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(out_n, (unsigned long long)__popc(m));
        basei = __shfl_sync(FULL, basei, 0);
        if (keep) {
            const uint64_t o = basei + __popc(m & ((1u << lane) - 1u));
            if (out_keys) {
                unsigned long long pk = planes_to_packed(s.x, s.y, k);        // (reads bits 0..k-1 of each plane: the tag is ignored)
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
// Conservation checks (tests, bench): sum of all counts in a table; number of valid k-mer windows of a record buffer.
// The second one shares nothing with the counting kernels (one thread per window, straight from the ASCII), so
// "sum of counts == valid windows" is an independent end-to-end check at any size and on any number of GPUs.
// =========================================================================================================
__global__ void __launch_bounds__(256)
k_table_sum(const Slot* __restrict__ slots, uint64_t cap, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = __ldcs(reinterpret_cast<const uint4*>(&slots[i]));
        if (slot_live(s, i, cap)) acc += s.z;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(256)
k_valid_windows(const uint8_t* __restrict__ recs, uint64_t nbytes, int k, unsigned long long* __restrict__ out) {
    // thread t owns positions [64 t, 64 t + 64): it walks 64 + k - 1 bytes keeping the length of the current run of bases
    unsigned long long acc = 0;
    const uint64_t nstrips = (nbytes + 63) / 64;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < nstrips; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = t * 64, e = min(nbytes, a + 64 + (uint64_t)k - 1);
        unsigned run = 0;
        for (uint64_t i = a; i < e; i++) {
            run = base_valid(recs[i]) ? run + 1u : 0u;
            // a window ENDS at i and starts at i - k + 1: counted by the strip that owns its start
            if (run >= (unsigned)k && i + 1 >= a + (uint64_t)k && i + 1 - k < a + 64) acc++;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

cudaError_t launch_table_sum(const Slot* slots, uint64_t cap, unsigned long long* d_out, cudaStream_t s) {
    TimedLaunch timed("k_table_sum", s);
    if (cap == 0) return cudaSuccess;
    k_table_sum<<<148 * 8, 256, 0, s>>>(slots, cap, d_out);
    return cudaGetLastError();
}
cudaError_t launch_valid_windows(const uint8_t* d_recs, uint64_t nbytes, int k, unsigned long long* d_out, cudaStream_t s) {
    TimedLaunch timed("k_valid_windows", s);
    if (nbytes == 0) return cudaSuccess;
    k_valid_windows<<<148 * 8, 256, 0, s>>>(d_recs, nbytes, k, d_out);
    return cudaGetLastError();
}

// =========================================================================================================
// GUPS: the measured random-access roofline for this table geometry
// =========================================================================================================
// Every thread keeps GUPS_FLIGHT independent accesses in flight (a dependent chain per thread would measure latency, not
// the memory system): mode 0 = 256-bit loads of two adjacent slots (one 32-B sector, what a lookup of the key-hashed walk
// issues), 1 = 8-B load + RED.add on the same slot (steady-state count), 2 = CAS + RED.add.
constexpr int GUPS_FLIGHT = 8;
template <int MODE>
__global__ void __launch_bounds__(256)
k_gups(Slot* slots, uint64_t cap, uint64_t nops, unsigned long long* sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (uint64_t i = tid; i < nops; i += GUPS_FLIGHT * stride) {
        unsigned long long idx[GUPS_FLIGHT];
#pragma unroll
        for (int u = 0; u < GUPS_FLIGHT; u++)
            idx[u] = __umul64hi(mix64(0x9E3779B97F4A7C15ull * (i + u * stride + 1)), cap / 2) * 2;     // a sector-aligned pair
        if (MODE == 0) {
            unsigned long long k0[GUPS_FLIGHT], w0[GUPS_FLIGHT], k1[GUPS_FLIGHT], w1[GUPS_FLIGHT];
#pragma unroll
            for (int u = 0; u < GUPS_FLIGHT; u++) {
                k0[u] = w0[u] = k1[u] = w1[u] = 0ull;
                if (i + u * stride < nops) ld_slot_pair(&slots[idx[u]], k0[u], w0[u], k1[u], w1[u]);
            }
#pragma unroll
            for (int u = 0; u < GUPS_FLIGHT; u++) acc += k0[u] + w0[u] + k1[u] + w1[u];
        } else {
            unsigned long long kv[GUPS_FLIGHT];
#pragma unroll
            for (int u = 0; u < GUPS_FLIGHT; u++) {
                kv[u] = 0;
                if (i + u * stride < nops) {
                    if (MODE == 1) kv[u] = __ldcg(&slots[idx[u]].key);
                    else kv[u] = atomicCAS(&slots[idx[u]].key, 0ull, KEY_TAG | idx[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < GUPS_FLIGHT; u++)
                if (i + u * stride < nops) { atomicAdd(&slots[idx[u]].val, 1u); acc += kv[u]; }
        }
    }
    if (acc == 0x123456789ull) *sink = acc;   // keep the loads alive
}

cudaError_t launch_gups(Slot* slots, uint64_t cap, uint64_t nops, int mode, unsigned long long* d_sink, int sm_count,
                        cudaStream_t s) {
    TimedLaunch timed("k_gups", s);
    const int grid = sm_count * 6;
    if (mode == 0) k_gups<0><<<grid, 256, 0, s>>>(slots, cap, nops, d_sink);
    else if (mode == 1) k_gups<1><<<grid, 256, 0, s>>>(slots, cap, nops, d_sink);
    else k_gups<2><<<grid, 256, 0, s>>>(slots, cap, nops, d_sink);
    return cudaGetLastError();
}

}  // namespace tg
