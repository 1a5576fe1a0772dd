// Internal launch interface between the C-ABI host layer (tg_api.cu) and the kernels (tg_kernels.cu,
// tg_sort.cu, tg_synth.cu).  Nothing here is exported.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <vector>
#include "tg_device.cuh"

namespace tg {

// ---- geometry of the flat-tile kernels (count / label) ------------------------------------------------
constexpr int CT_THREADS = 256;
constexpr int CT_TILE = CT_THREADS * 32;   // bases per tile: one 32-base chunk per thread
constexpr int CT_HALO = 32;                // one extra chunk so windows may run past the tile end (k <= 32)
constexpr int CT_LOAD = CT_TILE + CT_HALO; // bytes per TMA bulk copy (multiple of 16)
constexpr int REC_PAD = 64;                // '\n' bytes behind the last tile of a device record buffer: the weldmer kernel's
                                           // windows (up to 48 bases) read two chunks past a tile

// bytes a device record buffer of nbytes must be allocated (and '\n'-padded) to
inline uint64_t padded_record_bytes(uint64_t nbytes) {
    uint64_t ntiles = (nbytes + CT_TILE - 1) / CT_TILE;
    if (ntiles == 0) ntiles = 1;
    return ntiles * CT_TILE + REC_PAD;
}

// ---- geometry of the per-read kernels (stats / assign) ------------------------------------------------
constexpr int PR_WARPS = 8;          // warps (= reads in flight) per CTA
constexpr int PR_MAXWIN = 256;       // windows per read handled by the warp path (reads up to 256+k-1 bases)
constexpr int PR_MAXCH = ((PR_MAXWIN + 32) / 32 + 2 + 3) / 4 * 4;  // plane chunks per read incl. sentinel (whole groups of 4)
constexpr int LONG_THREADS = 256;    // CTA-per-read path for longer reads

struct LongList {            // filled by the warp-path kernels, consumed by the CTA-per-read kernels
    unsigned int* count;     // number of long reads found
    unsigned int* max_win;   // largest window count among them
    unsigned int* idx;       // their read indices (capacity = nreads)
};

int max_resident_ctas(const void* kernel, int threads, size_t dyn_smem, int device);

// ---- optional per-kernel timing (tg_ctx_set "kernel_timing" 1; read back with tg_kernel_times) -------------
// Every launch_* wrapper brackets its kernel with two CUDA events on the launching stream while a timer is bound
// and switched on.  Off (the default) it costs one branch.
struct KernelSpan { const char* name; cudaEvent_t a, b; };
struct KernelTimer { bool on = false; std::vector<KernelSpan> spans; };
extern thread_local KernelTimer* g_kernel_timer;
struct TimedLaunch {
    cudaStream_t s; long idx = -1;
    TimedLaunch(const char* name, cudaStream_t stream) : s(stream) {
        KernelTimer* kt = g_kernel_timer;
        if (!kt || !kt->on) return;
        KernelSpan sp{name, nullptr, nullptr};
        if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
        cudaEventRecord(sp.a, s);
        kt->spans.push_back(sp);
        idx = (long)kt->spans.size() - 1;
    }
    ~TimedLaunch() { if (idx >= 0) cudaEventRecord(g_kernel_timer->spans[idx].b, s); }
};

// flat tiles: count every valid k-mer window of a '\n'-padded record buffer into a count table
cudaError_t launch_count_tiles(const uint8_t* d_recs, uint64_t nbytes, int k, int canonical, TableView t,
                               int sm_count, cudaStream_t s);
// flat tiles: label every valid forward window with (bundle index + 1), highest index wins
// (d_offs are offsets in the caller's frame; d_recs[0] is the byte at offset rec_base of that frame)
cudaError_t launch_label_tiles(const uint8_t* d_recs, uint64_t nbytes, const uint64_t* d_offs, uint64_t rec_base,
                               uint64_t nbundles, uint32_t first_bundle_index, int k, TableView t, int sm_count,
                               cudaStream_t s);
// partitioned count path, phase 1: every valid window's key appended to the log bin of its hash partition
// (t.slots may be null: then a full bin raises lg.error instead of counting directly)
constexpr unsigned LOG_MAX_BINS = 8192;
constexpr unsigned LOG_CAP_ALIGN = 16;                 // log bin capacity granularity (= entries per chunk)
constexpr unsigned LOG_CAP_MAX = 0xF0000000u;           // per-bin cursors are 32-bit and may overshoot the capacity
cudaError_t launch_log_tiles(const uint8_t* d_recs, uint64_t nbytes, int k, int canonical, LogView lg, TableView t,
                             int sm_count, cudaStream_t s);
// phase 2: replay log segments [nsrc][nlocal][cap] (cursor [nsrc][nlocal]) bin-major into the table; the bins are
// global bins bin0..bin0+nlocal-1 of nbins_global, `groups` bins open at a time.  d_chunk_start: scratch of
// log_replay_plan_words(nsrc, nlocal, groups) u64.
size_t log_replay_plan_words(unsigned nsrc, unsigned nlocal, unsigned groups);
cudaError_t launch_log_replay(const LogEntry* d_keys, const unsigned int* d_cursor, unsigned cap, unsigned nsrc,
                              unsigned nlocal, unsigned bin0, unsigned nbins_global, unsigned groups,
                              unsigned long long* d_chunk_start, unsigned long long* d_hpoly, TableView t, int prefetch,
                              int sm_count, cudaStream_t s);
// owner-side refinement of a received COARSE log [nsrc][ncoarse][cap] into a FINE log [nfine][out_cap] (one segment per
// fine bin = table partition fine0 + bin of nfine_global): see k_log_refine.  d_out_cursor zeroed by the caller;
// d_chunk_start: scratch of log_refine_plan_words(nsrc, ncoarse) u64.  A full fine bin counts its overflow directly into
// overflow_table when that has slots, else raises error 3 in *d_error; a foreign key raises error 2.
size_t log_refine_plan_words(unsigned nsrc, unsigned ncoarse);
unsigned log_refine_max_split();      // most table partitions one coarse bin may cover
cudaError_t launch_log_refine(const LogEntry* d_keys, const unsigned int* d_cursor, unsigned cap, unsigned nsrc,
                              unsigned ncoarse, unsigned long long* d_chunk_start, LogEntry* d_out_keys,
                              unsigned int* d_out_cursor, unsigned out_cap, unsigned nfine, unsigned fine0,
                              unsigned nfine_global, int* d_error, TableView overflow_table, int sm_count, cudaStream_t s);
// (packed key, value) pairs -> table[canon(key)] += value (count tables) / max= (label tables)
cudaError_t launch_load_pairs(const uint64_t* d_keys, const uint32_t* d_vals, uint64_t n, int k, int canonical,
                              TableView t, int is_label, cudaStream_t s);
// re-insert every live slot of `from` whose value is in [min_val, max_val] into `to` (label tables: every slot)
cudaError_t launch_rehash(const Slot* from, uint64_t from_cap, TableView to, int is_label, uint32_t min_val,
                          uint32_t max_val, cudaStream_t s);

// per-read coverage statistics; offs are absolute offsets into the host buffer, rec_base is the offset of d_recs[0]
cudaError_t launch_cov_stats(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                             int canonical, const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean,
                             float* d_stdev, uint32_t* d_per_kmer, LongList ll, const uint32_t* d_order, int arena,
                             const uint32_t* d_counts /* nullable: counts per window position, looked up elsewhere */, cudaStream_t s);
cudaError_t launch_cov_stats_long(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k, int canonical,
                                  const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean, float* d_stdev,
                                  uint32_t* d_per_kmer, const unsigned int* d_long_idx, unsigned int n_long,
                                  unsigned int max_win, void* d_scratch, int nctas, cudaStream_t s);
size_t cov_stats_long_scratch_bytes(unsigned int max_win, int k, int nctas);
// device-driven twin: launched unconditionally, reads {count, max_win} from ll itself, lays out the given scratch budget
// (error 4 in *d_error when one read does not fit), returns at once when there is no long read.  No host sync.
cudaError_t launch_cov_stats_long_auto(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k,
                                       int canonical, const Slot* slots, Geo geo, uint32_t* d_median, float* d_mean,
                                       float* d_stdev, uint32_t* d_per_kmer, LongList ll, void* d_scratch,
                                       size_t scratch_bytes, int* d_error, int nctas, cudaStream_t s, const uint32_t* d_counts = nullptr);

cudaError_t launch_assign(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                          int strand, const Slot* slots, Geo geo, const uint8_t* d_entropy_ok,
                          int32_t* d_best, int32_t* d_pct, int32_t* d_score, LongList ll, const uint32_t* d_order, cudaStream_t s);
cudaError_t launch_assign_long(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k, int strand,
                               const Slot* slots, Geo geo, const uint8_t* d_entropy_ok, int32_t* d_best,
                               int32_t* d_pct, int32_t* d_score, const unsigned int* d_long_idx, unsigned int n_long,
                               unsigned int max_win, void* d_scratch, int nctas, cudaStream_t s);
size_t assign_long_scratch_bytes(unsigned int max_win, int k, int nctas);
cudaError_t launch_assign_long_auto(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, int k, int strand,
                                    const Slot* slots, Geo geo, const uint8_t* d_entropy_ok, int32_t* d_best,
                                    int32_t* d_pct, int32_t* d_score, LongList ll, void* d_scratch, size_t scratch_bytes,
                                    int* d_error, int nctas, cudaStream_t s);

// counting read by read, in the given order (tg_perread.cu): one ld.cg + RED.ADD per window straight into the table
cudaError_t launch_count_reads(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int k,
                               int canonical, TableView t, const uint32_t* d_order, cudaStream_t s);
// routed lookups (multi-GPU statistics without a replica): owner side and the way back (tg_kernels.cu)
cudaError_t launch_query_answer(const unsigned long long* d_keys, const unsigned int* d_cursor, unsigned cap, unsigned nsrc,
                                unsigned lp, const Slot* slots, Geo geo, unsigned int* d_resp, int sm_count, cudaStream_t s);
cudaError_t launch_query_scatter(const unsigned int* d_resp, const unsigned int* d_posidx, const unsigned int* d_cursor,
                                 unsigned nbins, unsigned cap, unsigned int* d_cov, int sm_count, cudaStream_t s);
// locus order of the reads (tg_perread.cu, tg_sort.cu): signature per read, then a radix sort of (signature, index)
cudaError_t launch_read_locus(const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads, int m,
                              uint32_t* d_sig, uint32_t* d_idx, int sm_count, cudaStream_t s);
size_t locus_sort_bytes(uint64_t n);
cudaError_t locus_sort(void* work, size_t work_bytes, uint64_t n, const uint32_t** d_sorted, cudaStream_t s);

// ---- GraphFromFasta weldmer counting: a read-only set of kk-mers (33 <= kk <= 48), every FORWARD
// window of every read that equals one of them bumps its counter (NonRedKmerTable::AddData, NonRedKmerTable.cc:162-200)
struct __align__(16) WeldSlot {
    unsigned long long lo;    // plane0 (kk bits) | low 16 bits of plane1 << 48
    unsigned int hi;          // plane1 >> 16
    unsigned int cnt;         // WELD_OCCUPIED | occurrences
};
constexpr unsigned WELD_OCCUPIED = 0x80000000u;
cudaError_t launch_weld_tiles(const uint8_t* d_recs, uint64_t nbytes, int kk, WeldSlot* slots, uint64_t cap_mask, int sm_count,
                              cudaStream_t s);
// table scans
cudaError_t launch_histo(const Slot* slots, uint64_t cap, unsigned long long* d_bins /*10002*/, cudaStream_t s);
cudaError_t launch_export(const Slot* slots, uint64_t cap, uint32_t min_count, uint32_t max_count, int k,
                          int canonical_repr, uint64_t* d_keys, uint32_t* d_vals, unsigned long long* d_n,
                          cudaStream_t s);
// conservation checks: sum of the counts of a table; valid k-mer windows of a record buffer (independent of the count path)
cudaError_t launch_table_sum(const Slot* slots, uint64_t cap, unsigned long long* d_out, cudaStream_t s);
cudaError_t launch_valid_windows(const uint8_t* d_recs, uint64_t nbytes, int k, unsigned long long* d_out, cudaStream_t s);
// tg_sort.cu: in-place ascending sort of (key,value) pairs on the low 2k bits (CUB radix sort)
cudaError_t sort_pairs(uint64_t* d_keys, uint32_t* d_vals, uint64_t n, int k, cudaStream_t s);

// random-access roofline probes (GUPS): mode 0 = 16-B loads, 1 = 8-B load + RED.add, 2 = CAS + RED.add
cudaError_t launch_gups(Slot* slots, uint64_t cap, uint64_t nops, int mode, unsigned long long* d_sink,
                        int sm_count, cudaStream_t s);

// tg_synth.cu: synthetic RNA-seq style reads generated on the device (bench / tests only)
cudaError_t launch_synth_reads(const uint8_t* d_tx, const uint64_t* d_tx_offs, const uint64_t* d_tx_cum,
                               uint32_t ntx, uint64_t npairs, int read_len, int frag_mean, int frag_sd,
                               uint32_t err_per_million, uint32_t n_per_million, uint64_t seed, int stranded,
                               uint8_t* d_recs, cudaStream_t s);

}  // namespace tg
