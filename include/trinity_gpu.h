/* trinity_gpu.h -- C ABI of libtrinity_gpu: the B200 (sm_100a) implementation of Trinity's k-mer hot path.
 *
 * The reference (trinityrnaseq v2.15.2) has NO in-process API for this path: its boundary is three
 * executables driven by shell strings (PerlLib/Pipeliner.pm:176, util/insilico_read_normalization.pl:803).
 * Each entry point below therefore names the reference *function* whose work it replaces; the three
 * drop-in executables (`jellyfish`, `fastaToKmerCoverageStats`, `ReadsToTranscripts`, built from
 * trinityrnaseq_b200/host/) are thin argv/file-format shells over these calls.  See INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every function returns TG_OK (0) or a negative TG_ERR_* code and leaves a
 *     message for tg_last_error() (thread-local).  There is no CPU fallback: without a usable CUDA device
 *     tg_init fails with TG_ERR_NOGPU.
 *   - "record buffer": sequences stored back to back, each followed by ONE terminator byte '\n'
 *     (exactly what a single-line FASTA sequence line looks like).  offs[i] is the byte offset of record i,
 *     offs[n] the end of the last terminator; record i has offs[i+1]-offs[i]-1 bases.  Bases are
 *     case-insensitive; any byte other than ACGTacgt is "not a base" and breaks every k-mer window over it.
 *   - "packed k-mer": 2 bits per base, A=0 C=1 G=2 T=3, first base most significant, in the low 2k bits of
 *     a uint64_t (so integer order == lexicographic order).  Count tables: 1 <= k <= 32 (Inchworm's range,
 *     Inchworm/src/KmerCounter.cpp:15-17; k = 32 is exact but runs on the direct-insert / CTA-per-read kernels and cannot
 *     be sharded); label tables and the partitioned / sharded paths: k <= 31.
 *   - host buffers may be pageable; buffers from tg_host_alloc (pinned) are copied at full PCIe speed.
 *   - a tg_ctx owns one device and its streams; calls on one ctx must not overlap in time.
 */
#ifndef TRINITY_GPU_H
#define TRINITY_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tg_ctx tg_ctx;
typedef struct tg_table tg_table;

enum { TG_OK = 0, TG_ERR_CUDA = -1, TG_ERR_ARG = -2, TG_ERR_NOMEM = -3, TG_ERR_TABLE = -4, TG_ERR_NOGPU = -5 };
enum { TG_TABLE_COUNT = 0, TG_TABLE_LABEL = 1 };
#define TG_HISTO_BINS 10002 /* bins[c] for c = 0..10000, bins[10001] = all larger counts */

/* ---- lifecycle -------------------------------------------------------------------------------------- */
int tg_version(void);
const char* tg_last_error(void);
int tg_device_count(void);
int tg_init(int device, tg_ctx** ctx);
void tg_destroy(tg_ctx* ctx);
int tg_device_info(tg_ctx* ctx, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes);
int tg_sync(tg_ctx* ctx);
/* tg_sync for the table-less log appends of the sharded count (tg_count_partition[_peers]_dev, tg_log_refine_dev): a full
 * bin -- a k-mer far hotter than the per-bin head-room allowed for -- is not an error but *overflowed = 1
 * (flag cleared): entries were dropped, so the caller repeats the batch with a larger per-bin capacity.  Every other
 * pending error is returned as by tg_sync. */
int tg_log_overflow_check(tg_ctx* ctx, int* overflowed);
/* number of kernels this ctx has launched so far (bench.py reports it as gpu_launches) */
uint64_t tg_launch_count(tg_ctx* ctx);

/* tuning knobs (also read from the environment at tg_init: TG_COUNT_MODE, TG_BATCH_MB, TG_PART_MB, TG_LOG_GB,
 * TG_REPLAY_PREFETCH): key = count_mode (auto|direct|log), batch_mb, batch_bytes, part_mb, part_bytes, log_gb,
 * log_bytes, replay_prefetch (0|1), replay_groups, replay_fold (0|1: fold a replay chunk's duplicate k-mers in shared
 * memory before the table; pays when several GPUs send their copies of the same k-mers to one owner), long_scratch_mb
 * (scratch budget of the device-resident entry points for reads beyond the warp path), kernel_timing (0|1).
 * None of them changes a result. */
int tg_ctx_set(tg_ctx* ctx, const char* key, const char* value);
/* With tg_ctx_set(ctx, "kernel_timing", "1") every kernel launch is bracketed by CUDA events on its stream;
 * tg_kernel_times syncs, writes one line "kernel-name \t total ms \t launches" per kernel into out, and resets. */
int tg_kernel_times(tg_ctx* ctx, char* out, uint64_t out_bytes);

void* tg_host_alloc(uint64_t bytes); /* pinned host memory */
void tg_host_free(void* p);
void tg_free(void* p); /* releases arrays returned by tg_table_export */

/* ---- k-mer tables ------------------------------------------------------------------------------------
 * TG_TABLE_COUNT replaces Inchworm's KmerCounter (hash_map<u64,u32>, Inchworm/src/KmerCounter.hpp:55) and
 * jellyfish's .jf hash; TG_TABLE_LABEL replaces Chrysalis' NonRedKmerTable (sorted vector<string> +
 * vector<int>, Chrysalis/analysis/NonRedKmerTable.h:84-86).  Tables grow on demand. */
int tg_table_create(tg_ctx* ctx, int kind, int k, uint64_t expected_keys, tg_table** out);
void tg_table_destroy(tg_table* t);
int tg_table_reserve(tg_table* t, uint64_t additional_keys);
int tg_table_info(tg_table* t, uint64_t* capacity_slots, uint64_t* distinct_keys);
int tg_table_clear(tg_table* t);     /* stream-ordered (see the device-resident section): no host synchronisation */
/* Geometry.  A table is `nparts` partitions of `slots_per_partition` slots.  A k-mer lives in the partition chosen by one
 * word of its hash and probes linearly, wrapping inside the partition, from the first slot of the 64-byte bucket (4 slots)
 * chosen by the other word.  slots_per_partition is rounded up to whole buckets.  tg_table_create picks nparts so that one partition fits in L2.  A SHARD holds the contiguous partition
 * range [part0, part0 + nlocal) of the global geometry -- the unit by which the table is split across GPUs
 * (owner(k-mer) = partition / nlocal; prior art: MPIinchworm's `canonical k-mer % NUM_MPI_NODES`,
 * Inchworm/src/mpi_deprecated/MPIinchworm.cpp:1236-1257).  The concatenation of all shards' slot arrays, in rank
 * order, is bit-for-bit the full table (part0 = 0, nlocal = nparts). */
int tg_table_create_sharded(tg_ctx* ctx, int kind, int k, uint64_t slots_per_partition, uint32_t nparts, uint32_t part0,
                            uint32_t nlocal, tg_table** out);
int tg_table_geometry(tg_table* t, uint64_t* slots_per_partition, uint32_t* nparts, uint32_t* part0, uint32_t* nlocal);
int tg_table_resize(tg_table* t, uint64_t slots_per_partition);      /* rehash into a new partition size */
/* `jellyfish dump -L min` kept on the device: the number of k-mers with count >= min_count, and a copy of exactly
 * those k-mers into `dst` (cleared first; same kind, k and partition range as t, any partition size).  This is the
 * table util/insilico_read_normalization.pl feeds to fastaToKmerCoverageStats (jellyfish dump -L 2, :45,641):
 * coverage statistics clamp every count below 1 to 1 (fastaToKmerCoverageStats.cpp:328-330), so a table without
 * its count-1 k-mers gives bit-identical statistics while being several times smaller. */
int tg_table_count_min(tg_table* t, uint32_t min_count, uint64_t* n);
int tg_table_compact_into(tg_table* t, uint32_t min_count, tg_table* dst);
/* raw slot array (16 B per slot) in HBM, for all-gathering shards into a full table, and the matching setter of
 * the distinct-key counter of a table assembled that way */
int tg_table_slots_dev(tg_table* t, void** d_slots, uint64_t* nbytes);
int tg_table_set_distinct(tg_table* t, uint64_t distinct);

/* ---- stage J: jellyfish count / dump / histo ----------------------------------------------------------
 * tg_count_reads: `jellyfish count -m k [--canonical]` (Trinity:2612-2619) and
 *   KmerCounter::add_sequence (Inchworm/src/KmerCounter.cpp:34-44): table[key(w)] += 1 for every window w of
 *   k bases in every record; canonical != 0 folds a k-mer and its reverse complement onto one key. */
int tg_count_reads(tg_table* t, const char* recs, uint64_t nbytes, int canonical);
/* tg_table_load_pairs: populate_kmer_counter_from_kmers / KmerCounter::add_kmer(kmer,count)
 *   (Inchworm/src/fastaToKmerCoverageStats.cpp:181-228, KmerCounter.cpp:476-489): table[canon(key)] += val
 *   (u32 wrap-around like the reference's unsigned int). */
int tg_table_load_pairs(tg_table* t, const uint64_t* packed_keys, const uint32_t* vals, uint64_t n, int canonical);
/* tg_table_export: `jellyfish dump -L min -U max` (Trinity:2625).  Returns malloc'ed arrays (tg_free) of
 *   packed k-mers and counts with min <= count <= max; sorted != 0 orders them by k-mer;
 *   canonical_repr != 0 reports the lexicographically smaller of k-mer / reverse complement (what jellyfish
 *   prints for a --canonical table). */
int tg_table_export(tg_table* t, uint32_t min_count, uint32_t max_count, int sorted, int canonical_repr,
                    uint64_t** packed_keys, uint32_t** counts, uint64_t* n);
/* tg_histo: `jellyfish histo` (Trinity:2630): bins[c] = number of distinct k-mers with count c. */
int tg_histo(tg_table* t, uint64_t bins[TG_HISTO_BINS]);
/* Conservation check used by the tests and the bench at sizes where no dump can be compared: the sum of all counts of a
 * count table must equal the number of valid k-mer windows of the reads counted into it.  tg_valid_windows_dev counts those
 * windows straight from the ASCII record buffer (device memory), sharing nothing with the counting kernels. */
int tg_table_count_sum(tg_table* t, uint64_t* sum);
int tg_valid_windows_dev(tg_ctx* ctx, const void* d_recs, uint64_t nbytes, int k, uint64_t* n);

/* ---- one upload for several calls ----------------------------------------------------------------------
 * The reference tools read the same reads twice when `--kmers_from_reads` names the `--reads` file (count, then
 * statistics: Inchworm/src/fastaToKmerCoverageStats.cpp:230-293, :122-172), and so do the two halves of a bench step.
 * tg_records_hold declares a host record buffer IMMUTABLE until tg_records_release (or the next hold): the library keeps
 * one device copy of it (uploaded during the first call that uses it, overlapped with that call's kernels) and every
 * later tg_count_reads / tg_cov_stats / tg_assign_reads given the same pointer and length (offs[nreads] for the per-read
 * calls) works on that copy.  Results are identical with or without it. */
int tg_records_hold(tg_ctx* ctx, const char* recs, uint64_t nbytes);
int tg_records_release(tg_ctx* ctx);

/* ---- stage S: fastaToKmerCoverageStats ------------------------------------------------------------------
 * compute_kmer_coverage + median_coverage + mean + stDev (Inchworm/src/fastaToKmerCoverageStats.cpp:300-402)
 * for every record: per window c = max(1, table[canon(w)]) (0 -> 1 also for windows with a non-base);
 * median as u32 (even n: wrapping u32 mean of the two middles), mean = (float)sum/n, stdev = sequential fp32
 * sqrtf(sum((c-mean)^2)/(n-1)) without FMA, bit-identical to the x86-64 reference (n==1 -> -nan, n==0 -> -0).
 * per_kmer (optional, may be NULL): per_kmer[offs[i]+j] = coverage of window j of record i
 * (--capture_coverage_info); it must have offs[nreads] entries. */
int tg_cov_stats(tg_table* t, const char* recs, const uint64_t* offs, uint64_t nreads, int canonical,
                 uint32_t* median, float* mean, float* stdev, uint32_t* per_kmer);

/* ---- stage R: ReadsToTranscripts ------------------------------------------------------------------------
 * tg_label_bundles: NonRedKmerTable::SetUp(dna,true) + the SetCount loop
 *   (Chrysalis/analysis/ReadsToTranscripts.cc:144-169): every all-ACGT forward k-mer of bundle i gets label
 *   first_index + i; a k-mer present in several bundles keeps the HIGHEST index (the reference's
 *   single-thread last-writer-wins order). */
int tg_label_bundles(tg_table* t, const char* recs, const uint64_t* offs, uint64_t nbundles, uint32_t first_index);
/* tg_assign_reads: the per-read loop of ReadsToTranscripts.cc:216-274.  entropy_ok is a 26*26*26 byte table
 *   indexed [nG][nA][nT] (tg_entropy_table builds it with the reference's expression); strand != 0 skips the
 *   reverse-complement pass.  Outputs per read: best = bundle index or -1, pct = pct_read_mapped,
 *   score (optional, may be NULL) = the winning run length `max`. */
int tg_assign_reads(tg_table* t, const char* recs, const uint64_t* offs, uint64_t nreads, int strand,
                    const uint8_t* entropy_ok, int32_t* best, int32_t* pct, int32_t* score);
/* compute_entropy(string&) >= min_entropy for every (nG,nA,nT,nC) with sum k
 * (Chrysalis/analysis/sequenceUtil.cc:326-355); host-side, evaluated with the reference's fp32 expression. */
void tg_entropy_table(int k, float min_entropy, uint8_t* entropy_ok /* 26*26*26 */);

/* ---- next row: GraphFromFasta weldmer counting ----------------------------------------
 * NonRedKmerTable::SetUp(crossover) + AddData(DNAStringStreamFast&) (Chrysalis/analysis/GraphFromFasta.cc:1412-1424,
 * NonRedKmerTable.cc:12-96,162-200): `weldmers` holds n candidate strings of kk characters back to back (33 <= kk <= 48; the
 * reference's default -kk is 48).  Every window of kk bases of every record that EQUALS a candidate -- forward strand only,
 * case-insensitive, no canonicalisation -- adds one to that candidate's counter.  tg_weld_counts returns the counters in
 * input order (duplicate candidates share one counter, like the reference's unique-sorted table; a candidate with a
 * non-ACGT character can match no read window and stays 0).  Counts are what GetCount(weldmer, 0) returns to the weld
 * decisions (GraphFromFasta.cc:538,584). */
typedef struct tg_weld tg_weld;
int tg_weld_create(tg_ctx* ctx, int kk, const char* weldmers, uint64_t n, tg_weld** out);
void tg_weld_destroy(tg_weld* w);
int tg_weld_count_reads(tg_weld* w, const char* recs, uint64_t nbytes);
int tg_weld_count_reads_dev(tg_weld* w, const void* d_recs, uint64_t nbytes);
int tg_weld_counts(tg_weld* w, int32_t* counts /* n */);

/* ---- device-resident variants (inputs already in HBM; used for kernel-only timing) ----------------------
 * These calls (and tg_table_clear) are STREAM-ORDERED and never synchronise with the host: they return once the work is
 * queued, results and error flags are valid after tg_sync (or any host-buffer call on the same table).  A caller can
 * therefore queue clear -> count -> statistics for several batches back to back; a stalled host thread then never
 * idles the GPU.  Reads with more than 256 windows are handled by a second kernel that is launched unconditionally and
 * sizes itself from device memory; a read too long for its scratch budget (tg_ctx_set "long_scratch_mb", default 64)
 * is reported by tg_sync -- the host-buffer entry points have no such limit. */
int tg_dev_alloc(tg_ctx* ctx, uint64_t bytes, void** dptr);
int tg_dev_records_alloc(tg_ctx* ctx, uint64_t nbytes, void** dptr); /* padded + '\n'-filled for the tile kernels */
int tg_dev_free(tg_ctx* ctx, void* dptr);
int tg_memcpy_h2d(tg_ctx* ctx, void* dptr, const void* host, uint64_t bytes);
int tg_memcpy_d2h(tg_ctx* ctx, void* host, const void* dptr, uint64_t bytes);
int tg_memcpy_d2d(tg_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int tg_memset_dev(tg_ctx* ctx, void* dst, int value, uint64_t bytes);
int tg_count_reads_dev(tg_table* t, const void* d_recs, uint64_t nbytes, int canonical);
/* Sharded counting, the two halves around the exchange (hash-sharded table across GPUs):
 *   tg_count_partition_dev  every k-mer occurrence of the record buffer appended to bin part(k-mer) of a
 *       caller-owned log: d_keys [nbins][cap] entries of tg_log_entry_bytes(), d_cursor [nbins] u32 (zeroed by the caller).  nbins is the
 *       table's global partition count or a divisor of it that is a multiple of the rank count (coarse bins, see
 *       tg_log_refine_dev), so bins [r*nbins/ranks, (r+1)*nbins/ranks) are exactly what rank r owns and one
 *       equal-split all-to-all of d_keys / d_cursor routes every k-mer to its owner.  A bin overflow is
 *       reported by the next tg_sync.  Homopolymer windows (poly-A tails...) bypass the log: d_hpoly is 8 u64,
 *       zeroed by the caller -- [0..3] table keys of A^k, C^k, G^k, T^k, [4..7] their occurrence counts; sum the
 *       counts (and max the keys) over ranks before the replay.
 *   tg_table_replay_log_dev  inserts a received log [nsrc][nlocal][cap] (+ cursors [nsrc][nlocal]) into the shard,
 *       plus the homopolymer tallies whose partition the shard holds (d_hpoly may be NULL; its counts are
 *       cleared). */
int tg_count_partition_dev(tg_ctx* ctx, const void* d_recs, uint64_t nbytes, int k, int canonical, uint32_t nbins,
                           uint32_t cap, void* d_keys, void* d_cursor, void* d_hpoly);
int tg_table_replay_log_dev(tg_table* t, const void* d_keys, const void* d_cursor, void* d_hpoly, uint32_t nsrc,
                            uint32_t cap);
/* Fused phase 1 + exchange (the path `north_star` asks for: k-mers routed to their owner GPU while counting).  One
 * process per GPU; every rank allocates its receive log [nranks][nbins/nranks][cap] u64 with tg_dev_alloc, exports it
 * with tg_ipc_export, and opens the other ranks' handles with tg_ipc_open (CUDA IPC; peer access over NVLink is
 * enabled lazily).  tg_count_partition_peers_dev is tg_count_partition_dev whose entries are stored straight into
 * segment `my_rank` of the OWNER's receive log (d_owner_keys[r] = rank r's log, the local pointer for r = my_rank):
 * the transfer overlaps the rolling of the next tile, no send buffer and no separate all-to-all of keys.  d_cursor
 * [nbins] stays local; after the call (and a tg_sync) the ranks exchange cursor rows -- rank r needs
 * cursor[r*lp .. (r+1)*lp) of every rank as its [nranks][lp] cursor array for tg_table_replay_log_dev -- which is also
 * the point after which every rank's stores have landed.  nbins/nranks must be a power of two, nranks <= 8.
 * Replaces the exchange the reference's retired MPI build did with one blocking MPI_Send per k-mer
 * (Inchworm/src/mpi_deprecated/MPIinchworm.cpp:519-531, owner rule :1236-1257). */
/* Coarse exchange bins, fine table partitions.  Phase 1 is fast only while a tile of reads scatters into a few hundred
 * bins (runs of consecutive entries per bin; see profiles/README.md), but a big sharded table has thousands of
 * L2-sized partitions.  So the ranks exchange COARSE bins -- nbins in the two partition calls above may be any
 * divisor of the table's partition count that is a multiple of the rank count -- and the owner splits what it received,
 * [nsrc][ncoarse][cap] with cursors [nsrc][ncoarse], into one segment per local partition: d_out_keys [nfine][out_cap],
 * d_out_cursor [nfine] (zeroed by the caller), partitions fine0 .. fine0+nfine-1 of nfine_global.  Then
 * tg_table_replay_log_dev(t, d_out_keys, d_out_cursor, hpoly, 1, out_cap).  Errors surface at the next tg_sync. */
int tg_log_refine_dev(tg_ctx* ctx, const void* d_keys, const void* d_cursor, uint32_t nsrc, uint32_t ncoarse, uint32_t cap,
                      void* d_out_keys, void* d_out_cursor, uint32_t nfine, uint32_t out_cap, uint32_t fine0,
                      uint32_t nfine_global);
/* bytes of one k-mer log entry (16: the table key and the packed home of the k-mer): every `d_keys` / `d_owner_keys` /
 * `d_out_keys` log of the sharded-count calls holds bins * cap entries of this size */
uint32_t tg_log_entry_bytes(void);
#define TG_IPC_HANDLE_BYTES 64
int tg_ipc_export(tg_ctx* ctx, void* dptr, uint8_t* handle /* TG_IPC_HANDLE_BYTES */);
int tg_ipc_open(tg_ctx* ctx, const uint8_t* handle, void** dptr);
int tg_ipc_close(tg_ctx* ctx, void* dptr);
int tg_count_partition_peers_dev(tg_ctx* ctx, const void* d_recs, uint64_t nbytes, int k, int canonical, uint32_t nbins,
                                 uint32_t cap, uint32_t nranks, uint32_t my_rank, void* const* d_owner_keys,
                                 void* d_cursor, void* d_hpoly);
/* `jellyfish dump -L n` as a VIEW: from now on tg_cov_stats[_dev] on this count table treats every k-mer whose count is
 * below min_count as absent -- exactly what the statistics of a table rebuilt from `dump -L min_count` are (the
 * normalisation pipeline's -L 2, util/insilico_read_normalization.pl:45,641), without materialising that table.  Counting,
 * dump, histo and export are unaffected.  0 or 1 = the whole table. */
int tg_table_set_count_floor(tg_table* t, uint32_t min_count);
/* ---- routed lookups: statistics against a SHARDED table without a replica (keys out, counts back, 12 B per
 * lookup; prior art for owner routing: Inchworm/src/mpi_deprecated/MPIinchworm.cpp:519-531,1236-1257) -------------------
 * requester: tg_query_partition_dev bins the key of every valid window by owner exactly like tg_count_partition_dev bins
 *   counted k-mers ([nbins][cap] log + cursors) and records, at the same (bin, pos), the window's position in the record
 *   buffer (d_posidx, u32 [nbins][cap]); the bins travel to their owners (all-to-all);
 * owner:     tg_query_answer_dev looks the received keys up in its shard -> d_resp, same layout, value or 0; the answers
 *   travel back;
 * requester: tg_query_scatter_dev puts every answer at its window's position (d_counts: u32 per byte of the record buffer,
 *   zeroed by the caller); tg_cov_stats_counts_dev computes the per-read statistics from those counts -- bit-identical
 *   to tg_cov_stats_dev on a replica (trinityrnaseq_b200/sharded.py: coverage_stats_routed_dev). */
int tg_query_partition_dev(tg_ctx* ctx, const void* d_recs, uint64_t nbytes, int k, int canonical, uint32_t nbins, uint32_t cap,
                           void* d_keys, void* d_cursor, void* d_posidx);
int tg_query_answer_dev(tg_table* t, const void* d_keys, const void* d_cursor, uint32_t nsrc, uint32_t lp, uint32_t cap,
                        void* d_resp);
int tg_query_scatter_dev(tg_ctx* ctx, const void* d_resp, const void* d_posidx, const void* d_cursor, uint32_t nbins, uint32_t cap,
                         void* d_counts);
int tg_cov_stats_counts_dev(tg_ctx* ctx, const void* d_recs, const void* d_offs, uint64_t nreads, int k, uint32_t count_floor,
                            const void* d_counts, void* d_median, void* d_mean, void* d_stdev);
/* Counting read by read, with the read offsets known (same record buffer + offs as tg_cov_stats_dev): the reads are
 * visited in LOCUS order (neighbouring reads cover the same stretch of a transcript), so the slots their k-mers share stay
 * L2-resident while they are incremented and no k-mer log / partition replay is needed.  Same counts as tg_count_reads_dev. */
int tg_count_records_dev(tg_table* t, const void* d_recs, const void* d_offs, uint64_t nreads, int canonical);
/* Declares a device record buffer (+ its offsets) immutable until the next tg_records_pin_dev (NULLs: nothing pinned): the
 * locus order of its reads is then computed once and shared by every *_dev call that is given exactly these pointers. */
int tg_records_pin_dev(tg_ctx* ctx, const void* d_recs, const void* d_offs, uint64_t nreads);
/* Queue the locus order of the pinned buffer NOW, on the library's second stream, so that it is computed beside whatever
 * the first stream does next (the count); later *_dev calls on the pinned buffer wait for it on the device.
 * recompute != 0: drop a cached order first (a benchmark step that must pay for the order every time). */
int tg_locus_prepare_dev(tg_ctx* ctx, int k, int recompute);
int tg_cov_stats_dev(tg_table* t, const void* d_recs, const void* d_offs, uint64_t nreads, int canonical,
                     void* d_median, void* d_mean, void* d_stdev);
int tg_label_bundles_dev(tg_table* t, const void* d_recs, uint64_t nbytes, const void* d_offs, uint64_t nbundles,
                         uint32_t first_index);
int tg_assign_reads_dev(tg_table* t, const void* d_recs, const void* d_offs, uint64_t nreads, int strand,
                        const void* d_entropy_ok, void* d_best, void* d_pct);
/* CUDA-event timer on the ctx's primary stream (the stream every *_dev call launches on) */
int tg_timer_start(tg_ctx* ctx);
int tg_timer_stop(tg_ctx* ctx, float* ms);

/* ---- measurement utilities ------------------------------------------------------------------------------ */
/* random-access roofline probe on a table of `slots` 16-B slots: mode 0 = random 16-B loads, 1 = 8-B load +
 * red.add on the same slot (steady-state count), 2 = CAS + red.add.  Returns the best of `reps` timings. */
int tg_gups(tg_ctx* ctx, uint64_t slots, uint64_t nops, int mode, int reps, float* best_ms);
/* synthetic paired reads generated on the device into a record buffer of 2*npairs*(read_len+1) bytes
 * (allocate it with tg_dev_records_alloc).  tx/tx_offs/tx_cum are HOST arrays: transcript bases, offsets
 * (ntx+1) and cumulative expression thresholds scaled to 2^64 (ntx). */
int tg_synth_reads_dev(tg_ctx* ctx, const char* tx, const uint64_t* tx_offs, const uint64_t* tx_cum, uint32_t ntx,
                       uint64_t npairs, int read_len, int frag_mean, int frag_sd, uint32_t err_per_million,
                       uint32_t n_per_million, uint64_t seed, int stranded, void* d_recs);

#ifdef __cplusplus
}
#endif
#endif /* TRINITY_GPU_H */
