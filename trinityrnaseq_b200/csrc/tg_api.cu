// C-ABI host layer of libtrinity_gpu (declared in include/trinity_gpu.h).
// Owns device memory, streams and batching; all arithmetic happens in the kernels of tg_kernels.cu.
// There is deliberately no CPU path here: if CUDA is unusable every call fails with an error code.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/trinity_gpu.h"
#include "tg_internal.h"

using namespace tg;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(e__ == cudaErrorMemoryAllocation ? TG_ERR_NOMEM : TG_ERR_CUDA, "%s failed: %s (%s:%d)", \
                        #call, cudaGetErrorString(e__), __FILE__, __LINE__);                             \
    } while (0)

constexpr double MAX_LOAD = 0.70;      // grow before a batch could push occupancy past this
constexpr double TARGET_LOAD = 0.45;   // occupancy right after a growth
constexpr uint64_t MIN_SLOTS = 1u << 12;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct tg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint64_t launches = 0;
    size_t batch_bytes = 64ull << 20;
    // per-stream staging for the host-buffer entry points
    DevBuf recs[2], offs[2], out_a[2], out_b[2], out_c[2], per_kmer[2], long_idx[2], scratch;
    DevBuf lut;
    DevBuf locus[2];                                    // per stream: signatures, indices, sort scratch of the locus order
    // locus orders that outlive a call: of the held host buffer's device copy, and of a device buffer the caller pinned
    // (tg_records_pin_dev).  Both buffers are immutable by contract, so their order is computed once.
    struct LocusCache {
        const void* recs = nullptr; const void* offs_key = nullptr; uint64_t nreads = 0; int m = 0;
        DevBuf buf; const uint32_t* order = nullptr; cudaEvent_t ready = nullptr; bool valid = false;
        void forget() { valid = false; order = nullptr; }
    } locus_held, locus_pin;
    const void* pin_recs = nullptr; const void* pin_offs = nullptr; uint64_t pin_nreads = 0;
    bool locus_order = true;                            // per-read kernels visit the reads in locus order (tg_perread.cu)
    uint64_t locus_min_reads = 1ull << 15;              // ... when a launch has at least this many reads
    int locus_m = 16;                                   // signature = smallest hash over the read's m-mers, m = min(k, locus_m)
    int stats_arena = 1024;                             // coverage words per warp of k_cov_stats (shared memory vs L1)
    DevBuf est_scratch;                                 // scratch table + counter of estimate_log_distinct (kept: cudaFree would synchronise the device, uploads included)
    DevBuf long_scratch;                                // fixed budget of the device-driven CTA-per-read kernels (*_dev entry points)
    size_t long_scratch_bytes = 64ull << 20;            // reads up to ~2 M windows; longer ones: host-buffer entry points
    cudaEvent_t order[2] = {nullptr, nullptr};          // cross-stream ordering without host syncs
    unsigned int* d_long_hdr[2] = {nullptr, nullptr};   // {count, max_win}
    unsigned int* h_long_hdr = nullptr;                 // pinned, 2 x 2
    int* d_error = nullptr;                             // raised by table-less log appends (tg_count_partition_dev)
    size_t part_bytes = 16ull << 20;                    // target bytes of one table partition (L2-resident unit)
    size_t log_max_bytes = 48ull << 30;                 // most HBM the k-mer log of one table may take
    // a host record buffer the caller declared immutable (tg_records_hold): its ONE device copy, shared by every call that
    // is given the same pointer and length
    struct Held {
        const char* host = nullptr; uint64_t nbytes = 0;
        DevBuf dev; bool uploaded = false;
        DevBuf offs, out_a, out_b, out_c, long_idx;
        std::vector<cudaEvent_t> chunk_events;             // one per upload chunk
    } held;
    size_t held_chunk_bytes = 64ull << 20;              // upload granularity of a held buffer (whole tiles)
    KernelTimer timer;                                  // optional per-kernel event timing
    int count_mode = 0;                                 // 0 auto, 1 always direct, 2 always logged
    int replay_prefetch = 1;
    bool replay_fold = false;                           // fold a chunk's duplicate keys in shared memory before the table
                                                        // (pays when many GPUs send their copies of the same hot k-mers to
                                                        // one owner; costs ~12 ms per 1.5 G entries otherwise)
    unsigned replay_groups = 8;                         // bins replayed concurrently (see k_log_replay)
};

// k-mer log of a count table (partitioned count path)
struct KeyLog {
    LogEntry* keys = nullptr;                  // [nbins][cap] entries
    unsigned int* cursor = nullptr;
    unsigned long long* chunk_start = nullptr;
    unsigned long long* hpoly = nullptr;   // [8] homopolymer side channel (keys, counts)
    unsigned nbins = 0, cap = 0;
    // tables with more partitions than LOG_BASE_BINS: phase 1 fills `nbins` COARSE bins (long runs per tile) and the
    // replay first splits them into one segment per partition (k_log_refine)
    LogEntry* fine_keys = nullptr;
    unsigned int* fine_cursor = nullptr;
    unsigned fine_bins = 0, fine_cap = 0, plan_bins = 0;
    uint64_t pending_ub = 0;          // host-side upper bound on entries appended since the last replay
    uint64_t total_entries() const { return (uint64_t)nbins * cap; }
};

struct tg_table {
    tg_ctx* ctx = nullptr;
    int kind = TG_TABLE_COUNT;
    int k = 25;
    Slot* slots = nullptr;
    Geo g{0, 1, 0, 1, 25, 0u};
    uint64_t cap = 0;                          // slots held here = g.nlocal * g.subcap
    unsigned long long* d_claimed = nullptr;   // [0] = distinct keys
    int* d_error = nullptr;
    uint64_t distinct_ub = 0;   // host-side upper bound on distinct keys (refreshed from the device at syncs)
    KeyLog log;
    bool sharded() const { return g.nlocal != g.nparts; }
    TableView view() const { return TableView{slots, g, d_claimed, d_error}; }
};

constexpr unsigned LOG_BASE_BINS = 512;

// partitions are whole 64-B buckets (tg_device.cuh: BUCKET_SLOTS)
static inline uint64_t whole_buckets(uint64_t slots) { return (slots + BUCKET_SLOTS - 1) / BUCKET_SLOTS * BUCKET_SLOTS; }

// partitions for a full table of `slots` slots: a power of two up to 512, beyond that multiples of 512, so that
// the log's bins (512 or the partition count) always nest with the partitions
static Geo pick_geo(uint64_t slots, size_t part_bytes) {
    const uint64_t bytes = slots * sizeof(Slot);
    uint64_t np = 1;
    while (np < LOG_BASE_BINS && bytes / np > part_bytes) np <<= 1;
    if (bytes / np > part_bytes) np = LOG_BASE_BINS * ((bytes + LOG_BASE_BINS * part_bytes - 1) / (LOG_BASE_BINS * part_bytes));
    if (np > LOG_MAX_BINS) np = LOG_MAX_BINS;
    Geo g;
    g.nparts = (unsigned)np; g.part0 = 0; g.nlocal = (unsigned)np;
    g.subcap = whole_buckets((slots + np - 1) / np);
    g.k = 0; g.floor = 0;            // set by table_new / table_regrow
    return g;
}

static int bind(tg_ctx* c) {
    g_kernel_timer = &c->timer;
    CU(cudaSetDevice(c->device));
    return TG_OK;
}

static int sync_all(tg_ctx* c) {
    CU(cudaStreamSynchronize(c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[1]));
    return TG_OK;
}

static int locus_order_async(tg_ctx* c, int b, const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads,
                             int k, const uint32_t** d_order, const void* held_offs_key = nullptr);

static int table_refresh(tg_table* t) {   // after a sync: read back distinct count and the error flag
    unsigned long long n = 0;
    int err = 0;
    CU(cudaMemcpy(&n, t->d_claimed, sizeof n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&err, t->d_error, sizeof err, cudaMemcpyDeviceToHost));
    t->distinct_ub = n;
    if (err == 2) return fail(TG_ERR_TABLE, "a k-mer was routed to a table shard that does not own its partition");
    if (err) return fail(TG_ERR_TABLE, "k-mer table overflow (capacity %llu slots in %u partitions)",
                         (unsigned long long)t->cap, t->g.nlocal);
    return TG_OK;
}

// Every table carries one spare bucket behind its partitions: its first slot is the home of the zero key (poly-A at
// k = 32, tg_device.cuh); the scan kernels read slot [cap] of any table.
constexpr uint64_t SPARE_SLOTS = BUCKET_SLOTS;
static int table_alloc(tg_ctx* c, uint64_t slots, Slot** out) {
    if (slots < MIN_SLOTS) slots = MIN_SLOTS;
    slots += SPARE_SLOTS;
    Slot* p = nullptr;
    CU(cudaMalloc(&p, slots * sizeof(Slot)));
    cudaError_t e = cudaMemsetAsync(p, 0, slots * sizeof(Slot), c->stream[0]);
    if (e != cudaSuccess) { cudaFree(p); CU(e); }
    *out = p;
    return TG_OK;
}

extern "C" {

int tg_version(void) { return 200; }
uint32_t tg_log_entry_bytes(void) { return (uint32_t)sizeof(LogEntry); }
const char* tg_last_error(void) { return g_err.c_str(); }

int tg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tg_init(int device, tg_ctx** out) {
    if (!out) return fail(TG_ERR_ARG, "tg_init: null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(TG_ERR_NOGPU, "no CUDA device available (%s); libtrinity_gpu has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(TG_ERR_ARG, "tg_init: device %d out of range (0..%d)", device, n - 1);
    tg_ctx* c = new tg_ctx();
    c->device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete c;
        return fail(TG_ERR_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    }
    c->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < 2; i++) {
        CU(cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
        CU(cudaMalloc(&c->d_long_hdr[i], 2 * sizeof(unsigned int)));
    }
    CU(cudaEventCreate(&c->t0));
    CU(cudaEventCreate(&c->t1));
    CU(cudaMallocHost(&c->h_long_hdr, 4 * sizeof(unsigned int)));
    CU(cudaMalloc(&c->d_error, sizeof(int)));
    CU(cudaMemset(c->d_error, 0, sizeof(int)));
    if (const char* mb = getenv("TG_BATCH_MB")) {
        long v = atol(mb);
        if (v >= 1 && v <= 4096) c->batch_bytes = (size_t)v << 20;
    }
    if (const char* mb = getenv("TG_PART_MB")) {          // bytes of one table partition (L2 blocking unit)
        long v = atol(mb);
        if (v >= 1 && v <= 1024) c->part_bytes = (size_t)v << 20;
    }
    if (const char* gb = getenv("TG_LOG_GB")) {           // HBM budget of the k-mer log
        long v = atol(gb);
        if (v >= 1 && v <= 160) c->log_max_bytes = (size_t)v << 30;
    }
    if (const char* m = getenv("TG_COUNT_MODE")) {        // auto | direct | log
        if (!strcmp(m, "direct")) c->count_mode = 1;
        else if (!strcmp(m, "log")) c->count_mode = 2;
    }
    if (const char* m = getenv("TG_REPLAY_PREFETCH")) c->replay_prefetch = atoi(m) != 0;
    *out = c;
    return TG_OK;
}

void tg_destroy(tg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) {
        c->recs[i].release(); c->offs[i].release(); c->out_a[i].release(); c->out_b[i].release();
        c->out_c[i].release(); c->per_kmer[i].release(); c->long_idx[i].release();
        if (c->d_long_hdr[i]) cudaFree(c->d_long_hdr[i]);
        if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
        if (c->done[i]) cudaEventDestroy(c->done[i]);
    }
    c->scratch.release(); c->lut.release(); c->long_scratch.release(); c->est_scratch.release(); c->locus[0].release(); c->locus[1].release(); c->locus_held.buf.release(); c->locus_pin.buf.release();
    if (c->locus_held.ready) cudaEventDestroy(c->locus_held.ready);
    if (c->locus_pin.ready) cudaEventDestroy(c->locus_pin.ready);
    c->held.dev.release(); c->held.offs.release(); c->held.out_a.release(); c->held.out_b.release(); c->held.out_c.release();
    c->held.long_idx.release();
    for (cudaEvent_t e : c->held.chunk_events) cudaEventDestroy(e);
    if (c->d_error) cudaFree(c->d_error);
    if (c->h_long_hdr) cudaFreeHost(c->h_long_hdr);
    for (int i = 0; i < 2; i++) if (c->order[i]) cudaEventDestroy(c->order[i]);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    delete c;
}

int tg_device_info(tg_ctx* c, int* sm_count, uint64_t* free_bytes, uint64_t* total_bytes) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = c->sm_count;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return TG_OK;
}

int tg_sync(tg_ctx* c) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    if (rc) return rc;
    int err = 0;
    CU(cudaMemcpy(&err, c->d_error, sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
        CU(cudaMemset(c->d_error, 0, sizeof(int)));
        if (err == 4) return fail(TG_ERR_ARG, "a read is too long for the scratch of the device-resident entry points "
                                              "(tg_cov_stats_dev / tg_assign_reads_dev): use the host-buffer entry points "
                                              "or raise long_scratch_mb");
        if (err == 2) return fail(TG_ERR_TABLE, "a k-mer reached a rank that does not own its partition (tg_log_refine_dev)");
        return fail(TG_ERR_TABLE, "a k-mer log bin overflowed (tg_count_partition_dev / tg_log_refine_dev): raise the per-bin capacity");
    }
    return TG_OK;
}

int tg_log_overflow_check(tg_ctx* c, int* overflowed) {
    if (!c || !overflowed) return fail(TG_ERR_ARG, "tg_log_overflow_check: null argument");
    *overflowed = 0;
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    if (rc) return rc;
    int err = 0;
    CU(cudaMemcpy(&err, c->d_error, sizeof err, cudaMemcpyDeviceToHost));
    if (err == 3) {                      // a bin was full: the caller lays the log out again and repeats the batch
        CU(cudaMemset(c->d_error, 0, sizeof(int)));
        *overflowed = 1;
        return TG_OK;
    }
    return tg_sync(c);                   // anything else is reported (and cleared) the usual way
}

uint64_t tg_launch_count(tg_ctx* c) { return c ? c->launches : 0; }

int tg_ctx_set(tg_ctx* c, const char* key, const char* value) {
    if (!c || !key || !value) return fail(TG_ERR_ARG, "tg_ctx_set: null argument");
    const long v = atol(value);
    if (!strcmp(key, "count_mode")) {
        if (!strcmp(value, "auto")) c->count_mode = 0;
        else if (!strcmp(value, "direct")) c->count_mode = 1;
        else if (!strcmp(value, "log")) c->count_mode = 2;
        else return fail(TG_ERR_ARG, "count_mode must be auto, direct or log");
    } else if (!strcmp(key, "batch_mb")) {
        if (v < 1 || v > 4096) return fail(TG_ERR_ARG, "batch_mb out of range");
        c->batch_bytes = (size_t)v << 20;
    } else if (!strcmp(key, "batch_bytes")) {
        if (v < 64) return fail(TG_ERR_ARG, "batch_bytes out of range");
        c->batch_bytes = (size_t)v;
    } else if (!strcmp(key, "part_mb")) {
        if (v < 1 || v > 1024) return fail(TG_ERR_ARG, "part_mb out of range");
        c->part_bytes = (size_t)v << 20;
    } else if (!strcmp(key, "part_bytes")) {
        if (v < 256) return fail(TG_ERR_ARG, "part_bytes out of range");
        c->part_bytes = (size_t)v;
    } else if (!strcmp(key, "log_gb")) {
        if (v < 1 || v > 160) return fail(TG_ERR_ARG, "log_gb out of range");
        c->log_max_bytes = (size_t)v << 30;
    } else if (!strcmp(key, "log_bytes")) {
        if (v < 4096) return fail(TG_ERR_ARG, "log_bytes out of range");
        c->log_max_bytes = (size_t)v;
    } else if (!strcmp(key, "replay_prefetch")) {
        c->replay_prefetch = v != 0;
    } else if (!strcmp(key, "long_scratch_mb")) {
        if (v < 1 || v > (64 << 10)) return fail(TG_ERR_ARG, "long_scratch_mb out of range (1..65536)");
        c->long_scratch_bytes = (size_t)v << 20;
    } else if (!strcmp(key, "replay_fold")) {
        c->replay_fold = v != 0;
    } else if (!strcmp(key, "replay_groups")) {
        if (v < 1 || v > 64) return fail(TG_ERR_ARG, "replay_groups out of range (1..64)");
        c->replay_groups = (unsigned)v;
    } else if (!strcmp(key, "locus_order")) {
        c->locus_order = v != 0;
    } else if (!strcmp(key, "locus_m")) {
        if (v < 8 || v > 31) return fail(TG_ERR_ARG, "locus_m out of range (8..31)");
        c->locus_m = (int)v;
    } else if (!strcmp(key, "stats_arena")) {
        if (v < 256 || v > 4096) return fail(TG_ERR_ARG, "stats_arena out of range (256..4096)");
        c->stats_arena = (int)v;
    } else if (!strcmp(key, "locus_min_reads")) {
        if (v < 0) return fail(TG_ERR_ARG, "locus_min_reads out of range");
        c->locus_min_reads = (uint64_t)v;
    } else if (!strcmp(key, "kernel_timing")) {
        c->timer.on = v != 0;
    } else {
        return fail(TG_ERR_ARG, "tg_ctx_set: unknown option '%s'", key);
    }
    return TG_OK;
}

void* tg_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        fail(TG_ERR_NOMEM, "cudaMallocHost(%llu) failed", (unsigned long long)bytes);
        return nullptr;
    }
    return p;
}
void tg_host_free(void* p) { if (p) cudaFreeHost(p); }
void tg_free(void* p) { free(p); }

// ---------------------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------------------
static int table_new(tg_ctx* c, int kind, int k, Geo g, tg_table** out) {
    tg_table* t = new tg_table();
    g.k = k; g.floor = 0u;
    if (g.subcap >= (1ull << 34)) return fail(TG_ERR_ARG, "a table partition holds at most 2^34 slots (bucket index is 32-bit)");
    t->ctx = c; t->kind = kind; t->k = k; t->g = g;
    t->cap = (uint64_t)g.nlocal * g.subcap;
    int rc = table_alloc(c, t->cap, &t->slots);
    if (rc) { delete t; return rc; }
    CU(cudaMalloc(&t->d_claimed, sizeof(unsigned long long)));
    CU(cudaMalloc(&t->d_error, sizeof(int)));
    CU(cudaMemsetAsync(t->d_claimed, 0, sizeof(unsigned long long), c->stream[0]));
    CU(cudaMemsetAsync(t->d_error, 0, sizeof(int), c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    *out = t;
    return TG_OK;
}

int tg_table_create(tg_ctx* c, int kind, int k, uint64_t expected_keys, tg_table** out) {
    if (!c || !out) return fail(TG_ERR_ARG, "tg_table_create: null argument");
    if (k < 1 || k > 32) return fail(TG_ERR_ARG, "k-mer length %d unsupported (1..32)", k);
    if (kind != TG_TABLE_COUNT && kind != TG_TABLE_LABEL) return fail(TG_ERR_ARG, "unknown table kind %d", kind);
    if (kind == TG_TABLE_LABEL && k > 31)        // (ReadsToTranscripts.cc:38 hard-codes k = 25; the entropy LUT holds k <= 25)
        return fail(TG_ERR_ARG, "label tables take k-mer lengths 1..31");
    if (bind(c)) return TG_ERR_CUDA;
    uint64_t slots = (uint64_t)((double)expected_keys / TARGET_LOAD) + 1;
    if (slots < MIN_SLOTS) slots = MIN_SLOTS;
    // expected_keys is a hint -- the executables derive it from a file size -- and the table grows on demand: never start
    // with more than the device can reasonably give (an input of hundreds of gigabytes would ask for more than HBM holds)
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
        const uint64_t fit = (uint64_t)((double)fr * 0.60) / sizeof(Slot);
        if (slots > fit && fit >= MIN_SLOTS) slots = fit;
    } else {
        cudaGetLastError();
    }
    return table_new(c, kind, k, pick_geo(slots, c->part_bytes), out);
}

int tg_table_create_sharded(tg_ctx* c, int kind, int k, uint64_t slots_per_partition, uint32_t nparts, uint32_t part0,
                            uint32_t nlocal, tg_table** out) {
    if (!c || !out) return fail(TG_ERR_ARG, "tg_table_create_sharded: null argument");
    if (k < 1 || k > 32) return fail(TG_ERR_ARG, "k-mer length %d unsupported (1..32)", k);
    if (k == 32 && nlocal != nparts)
        return fail(TG_ERR_ARG, "k = 32 tables cannot be sharded (the zero-key slot has no owner partition): 1..31");
    if (kind != TG_TABLE_COUNT && kind != TG_TABLE_LABEL) return fail(TG_ERR_ARG, "unknown table kind %d", kind);
    if (nparts == 0 || nlocal == 0 || (uint64_t)part0 + nlocal > nparts || nparts > LOG_MAX_BINS || slots_per_partition < 16)
        return fail(TG_ERR_ARG, "tg_table_create_sharded: bad geometry (%u partitions, local %u..+%u, %llu slots each)",
                    nparts, part0, nlocal, (unsigned long long)slots_per_partition);
    if (bind(c)) return TG_ERR_CUDA;
    Geo g;
    g.subcap = whole_buckets(slots_per_partition); g.nparts = nparts; g.part0 = part0; g.nlocal = nlocal; g.k = k;
    return table_new(c, kind, k, g, out);
}

int tg_table_geometry(tg_table* t, uint64_t* slots_per_partition, uint32_t* nparts, uint32_t* part0, uint32_t* nlocal) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    if (slots_per_partition) *slots_per_partition = t->g.subcap;
    if (nparts) *nparts = t->g.nparts;
    if (part0) *part0 = t->g.part0;
    if (nlocal) *nlocal = t->g.nlocal;
    return TG_OK;
}

static void log_release(tg_table* t) {
    if (t->log.keys) cudaFree(t->log.keys);
    if (t->log.cursor) cudaFree(t->log.cursor);
    if (t->log.chunk_start) cudaFree(t->log.chunk_start);
    if (t->log.fine_keys) cudaFree(t->log.fine_keys);
    if (t->log.fine_cursor) cudaFree(t->log.fine_cursor);
    if (t->log.hpoly) cudaFree(t->log.hpoly);
    t->log = KeyLog();
}

void tg_table_destroy(tg_table* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaDeviceSynchronize();
    log_release(t);
    if (t->slots) cudaFree(t->slots);
    if (t->d_claimed) cudaFree(t->d_claimed);
    if (t->d_error) cudaFree(t->d_error);
    delete t;
}

int tg_table_set_count_floor(tg_table* t, uint32_t min_count) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    if (t->kind != TG_TABLE_COUNT) return fail(TG_ERR_ARG, "tg_table_set_count_floor needs a TG_TABLE_COUNT table");
    t->g.floor = min_count;              // read by the kernels launched from now on (the geometry travels by value)
    return TG_OK;
}

int tg_table_clear(tg_table* t) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    // stream-ordered, no host synchronisation (a host stall must not idle the GPU): stream 0 waits for whatever stream 1
    // still does with the table, clears, and stream 1 waits for the clear
    if (!c->order[0]) {
        CU(cudaEventCreateWithFlags(&c->order[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->order[1], cudaEventDisableTiming));
    }
    CU(cudaEventRecord(c->order[1], c->stream[1]));
    CU(cudaStreamWaitEvent(c->stream[0], c->order[1], 0));
    CU(cudaMemsetAsync(t->slots, 0, (t->cap + SPARE_SLOTS) * sizeof(Slot), c->stream[0]));
    CU(cudaMemsetAsync(t->d_claimed, 0, sizeof(unsigned long long), c->stream[0]));
    CU(cudaMemsetAsync(t->d_error, 0, sizeof(int), c->stream[0]));
    if (t->log.cursor) {
        CU(cudaMemsetAsync(t->log.cursor, 0, t->log.nbins * sizeof(unsigned int), c->stream[0]));
        CU(cudaMemsetAsync(t->log.hpoly, 0, 8 * sizeof(unsigned long long), c->stream[0]));
    }
    t->log.pending_ub = 0;
    CU(cudaEventRecord(c->order[0], c->stream[0]));
    CU(cudaStreamWaitEvent(c->stream[1], c->order[0], 0));
    t->distinct_ub = 0;
    return TG_OK;
}

// Re-insert the k-mers of t with count >= min_count into `to` (growth, `dump -L n` on the device).
static int rehash_by_priority(tg_table* t, TableView to, uint32_t min_count) {
    tg_ctx* c = t->ctx;
    CU(launch_rehash(t->slots, t->cap, to, t->kind == TG_TABLE_LABEL, min_count, 0xFFFFFFFFu, c->stream[0]));
    c->launches++;
    return TG_OK;
}

// Move the table into a new geometry (growth).  Both streams must be idle.
static int table_regrow(tg_table* t, Geo ng) {
    tg_ctx* c = t->ctx;
    ng.k = t->k; ng.floor = t->g.floor;
    const uint64_t ncap = (uint64_t)ng.nlocal * ng.subcap;
    Slot* fresh = nullptr;
    int rc;
    if ((rc = table_alloc(c, ncap, &fresh))) return rc;
    CU(cudaMemsetAsync(t->d_claimed, 0, sizeof(unsigned long long), c->stream[0]));
    TableView nv{fresh, ng, t->d_claimed, t->d_error};
    if (int rc2 = rehash_by_priority(t, nv, 0)) { cudaFree(fresh); return rc2; }
    CU(cudaStreamSynchronize(c->stream[0]));
    CU(cudaFree(t->slots));
    t->slots = fresh;
    t->cap = ncap;
    t->g = ng;
    return table_refresh(t);
}

// geometry for a table that must hold `keys` distinct keys at TARGET_LOAD (shards keep their partition range)
static int grown_geo(tg_table* t, uint64_t keys, Geo* out) {
    tg_ctx* c = t->ctx;
    uint64_t want = (uint64_t)((double)keys / TARGET_LOAD) + 1;
    if (want < t->cap + t->cap / 2) want = t->cap + t->cap / 2;
    // never ask for more than the device can hold next to the old table
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    const uint64_t fit = (uint64_t)((double)fr * 0.92) / sizeof(Slot);
    if (want > fit) want = fit;
    if ((double)keys > 0.92 * (double)want)
        return fail(TG_ERR_NOMEM, "k-mer table cannot grow: need room for %llu keys, device has room for %llu slots",
                    (unsigned long long)keys, (unsigned long long)fit);
    if (t->sharded()) {
        *out = t->g;
        out->subcap = whole_buckets((want + t->g.nlocal - 1) / t->g.nlocal);
    } else {
        *out = pick_geo(want, c->part_bytes);
    }
    return TG_OK;
}

// Make room for `additional` more distinct keys.  distinct_ub is an upper bound; it is tightened from the
// device counter (one sync) only when the bound alone would force a growth.
int tg_table_reserve(tg_table* t, uint64_t additional) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if ((double)(t->distinct_ub + additional) <= MAX_LOAD * (double)t->cap) {
        t->distinct_ub += additional;
        return TG_OK;
    }
    int rc = sync_all(c);
    if (rc) return rc;
    if ((rc = table_refresh(t))) return rc;
    if ((double)(t->distinct_ub + additional) <= MAX_LOAD * (double)t->cap) {
        t->distinct_ub += additional;
        return TG_OK;
    }
    Geo ng;
    if ((rc = grown_geo(t, t->distinct_ub + additional, &ng))) return rc;
    if ((rc = table_regrow(t, ng))) return rc;
    t->distinct_ub += additional;
    return TG_OK;
}

int tg_table_resize(tg_table* t, uint64_t slots_per_partition) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    if (rc) return rc;
    if ((rc = table_refresh(t))) return rc;
    if ((double)t->distinct_ub > 0.92 * (double)slots_per_partition * t->g.nlocal)
        return fail(TG_ERR_ARG, "tg_table_resize: %llu keys do not fit %u x %llu slots", (unsigned long long)t->distinct_ub,
                    t->g.nlocal, (unsigned long long)slots_per_partition);
    Geo ng = t->g;
    ng.subcap = whole_buckets(slots_per_partition);
    return table_regrow(t, ng);
}

static int flush_log(tg_table* t, bool only_stream0 = false);

int tg_table_info(tg_table* t, uint64_t* capacity, uint64_t* distinct) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    if (bind(t->ctx)) return TG_ERR_CUDA;
    int rc = flush_log(t);
    if (rc) return rc;
    if ((rc = sync_all(t->ctx))) return rc;
    if ((rc = table_refresh(t))) return rc;
    if (capacity) *capacity = t->cap;
    if (distinct) *distinct = t->distinct_ub;
    return TG_OK;
}

// number of keys with value >= min_count (streams the table once)
static int count_min(tg_table* t, uint32_t min_count, uint64_t* n) {
    tg_ctx* c = t->ctx;
    unsigned long long* d_n = nullptr;
    CU(cudaMalloc(&d_n, sizeof *d_n));
    CU(cudaMemsetAsync(d_n, 0, sizeof *d_n, c->stream[0]));
    CU(launch_export(t->slots, t->cap, min_count, 0xFFFFFFFFu, t->k, 0, nullptr, nullptr, d_n, c->stream[0]));
    c->launches++;
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, d_n, sizeof v, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    cudaFree(d_n);
    *n = v;
    return TG_OK;
}

int tg_table_count_min(tg_table* t, uint32_t min_count, uint64_t* n) {
    if (!t || !n) return fail(TG_ERR_ARG, "tg_table_count_min: null argument");
    if (bind(t->ctx)) return TG_ERR_CUDA;
    int rc = flush_log(t);
    if (rc) return rc;
    if ((rc = sync_all(t->ctx))) return rc;
    return count_min(t, min_count, n);
}

int tg_table_compact_into(tg_table* t, uint32_t min_count, tg_table* dst) {
    if (!t || !dst || t == dst) return fail(TG_ERR_ARG, "tg_table_compact_into: bad argument");
    if (t->ctx != dst->ctx || t->kind != dst->kind || t->k != dst->k || t->g.nparts != dst->g.nparts ||
        t->g.part0 != dst->g.part0 || t->g.nlocal != dst->g.nlocal)
        return fail(TG_ERR_ARG, "tg_table_compact_into: the destination must have the source's kind, k and partition range");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc = flush_log(t);
    if (rc) return rc;
    if ((rc = tg_table_clear(dst))) return rc;       // syncs both streams
    if ((rc = rehash_by_priority(t, dst->view(), min_count))) return rc;
    return TG_OK;                                    // stream-ordered; an overfull destination raises its error flag
}

int tg_table_slots_dev(tg_table* t, void** d_slots, uint64_t* nbytes) {
    if (!t || !d_slots || !nbytes) return fail(TG_ERR_ARG, "tg_table_slots_dev: null argument");
    if (bind(t->ctx)) return TG_ERR_CUDA;
    int rc = flush_log(t);
    if (rc) return rc;
    if ((rc = sync_all(t->ctx))) return rc;
    *d_slots = t->slots;
    *nbytes = t->cap * sizeof(Slot);
    return TG_OK;
}

int tg_table_set_distinct(tg_table* t, uint64_t distinct) {
    if (!t) return fail(TG_ERR_ARG, "null table");
    if (bind(t->ctx)) return TG_ERR_CUDA;
    unsigned long long v = distinct;
    CU(cudaMemcpy(t->d_claimed, &v, sizeof v, cudaMemcpyHostToDevice));
    t->distinct_ub = distinct;
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// helpers for the host-buffer entry points
// ---------------------------------------------------------------------------------------------------------
// end (exclusive) of the batch starting at pos: at most batch_bytes, cut after a record terminator
static uint64_t batch_end(const char* recs, uint64_t pos, uint64_t nbytes, size_t batch_bytes) {
    uint64_t end = pos + batch_bytes;
    if (end >= nbytes) return nbytes;
    const void* nl = memrchr(recs + pos, '\n', end - pos);
    if (nl) return (const char*)nl - recs + 1;
    nl = memchr(recs + end, '\n', nbytes - end);   // one record longer than a batch: take all of it
    return nl ? (uint64_t)((const char*)nl - recs + 1) : nbytes;
}

static int upload_records(tg_ctx* c, int b, const char* src, uint64_t n) {
    const uint64_t padded = padded_record_bytes(n);
    CU(c->recs[b].ensure(padded));
    CU(cudaMemcpyAsync(c->recs[b].p, src, n, cudaMemcpyHostToDevice, c->stream[b]));
    CU(cudaMemsetAsync((char*)c->recs[b].p + n, '\n', padded - n, c->stream[b]));
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stage J
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"

// ---- k-mer log management (partitioned count path) -------------------------------------------------------
// can a log of `nbins` coarse bins be split into the table's partitions at replay time?
static bool log_refines(const tg_table* t, unsigned nbins) {
    const unsigned np = t->g.nparts;
    return !t->sharded() && nbins < np && np % nbins == 0 && np / nbins <= log_refine_max_split();
}

static unsigned log_bins_for(const tg_table* t) {
    if (t->sharded()) return t->g.nparts;                      // bins == global partitions (exchange unit)
    if (t->g.nparts <= LOG_BASE_BINS || log_refines(t, LOG_BASE_BINS)) return LOG_BASE_BINS;
    return t->g.nparts;
}

// worth logging?  The replay streams the whole table through L2 once, so the batch must be large next to it.
static bool log_pays(const tg_table* t, uint64_t nbytes) {
    const tg_ctx* c = t->ctx;
    if (t->kind != TG_TABLE_COUNT || t->sharded() || t->k < MIN_FAST_K || t->k > 31) return false;   // (k = 32: flat tiles only)
    if (c->count_mode == 1) return false;
    if (c->count_mode == 2) return true;
    return t->g.nparts >= 4 && nbytes * 16 >= t->cap * sizeof(Slot);
}

constexpr uint64_t LOG_BIN_SLACK = 1024;    // additive head-room per bin (hash fluctuation of small batches)
constexpr double LOG_BIN_FACTOR = 1.2;      // multiplicative head-room per bin (hot k-mers)

// log entries one phase-1 launch over nbytes record bytes can take up at most: one per byte
static uint64_t log_launch_cost(const tg_ctx*, const KeyLog&, uint64_t nbytes) { return nbytes; }

// entries that can be appended to an empty log without any bin expected to overflow
static uint64_t log_room(const KeyLog& lg) {
    if (lg.cap <= LOG_BIN_SLACK) return 0;
    return (uint64_t)((double)(lg.cap - LOG_BIN_SLACK) / LOG_BIN_FACTOR * lg.nbins) + lg.nbins;     // (+ rounding of cap)
}

// a log laid out for `entries` appended entries per replay (fewer if the HBM budget is smaller); *ok = false when
// no useful log fits
static int ensure_log(tg_table* t, uint64_t entries, bool* ok) {
    tg_ctx* c = t->ctx;
    *ok = false;
    const unsigned nbins = log_bins_for(t);
    const uint64_t want = ((uint64_t)((double)entries / nbins * LOG_BIN_FACTOR) + LOG_BIN_SLACK + LOG_CAP_ALIGN) / LOG_CAP_ALIGN * LOG_CAP_ALIGN;
    if (t->log.keys && t->log.nbins == nbins && t->log.cap >= std::min<uint64_t>(want, LOG_CAP_MAX) / LOG_CAP_ALIGN * LOG_CAP_ALIGN) {
        *ok = true;                                     // the common case costs no driver call (cudaMemGetInfo may block)
        return TG_OK;
    }
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    uint64_t budget = std::min<uint64_t>(c->log_max_bytes, fr / 2);
    if (log_refines(t, nbins)) budget /= 2;                  // the other half is the fine log of the replay
    if (t->log.keys) budget = std::max<uint64_t>(budget, t->log.total_entries() * sizeof(LogEntry));
    uint64_t per_bin = want;
    per_bin = std::min<uint64_t>(per_bin, budget / sizeof(LogEntry) / nbins);
    per_bin = std::min<uint64_t>(per_bin, LOG_CAP_MAX);
    per_bin = per_bin / LOG_CAP_ALIGN * LOG_CAP_ALIGN;
    if (per_bin < want && per_bin < 2 * LOG_BIN_SLACK) return TG_OK;    // budget-limited down to a useless size
    if (t->log.keys && t->log.nbins == nbins && t->log.cap >= per_bin) { *ok = true; return TG_OK; }
    if (t->log.pending_ub) { *ok = t->log.keys != nullptr; return TG_OK; }   // holds entries: keep its layout
    log_release(t);
    if (cudaMalloc(&t->log.keys, per_bin * nbins * sizeof(LogEntry)) != cudaSuccess) { cudaGetLastError(); t->log.keys = nullptr; return TG_OK; }
    CU(cudaMalloc(&t->log.cursor, nbins * sizeof(unsigned int)));
    CU(cudaMalloc(&t->log.chunk_start, log_replay_plan_words(1, nbins, 64) * sizeof(unsigned long long)));
    t->log.plan_bins = nbins;
    CU(cudaMalloc(&t->log.hpoly, 8 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(t->log.cursor, 0, nbins * sizeof(unsigned int), c->stream[0]));
    CU(cudaMemsetAsync(t->log.hpoly, 0, 8 * sizeof(unsigned long long), c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    t->log.nbins = nbins; t->log.cap = (unsigned)per_bin; t->log.pending_ub = 0;
    *ok = true;
    return TG_OK;
}

// a log held entirely by this GPU: one owner, one source segment
static LogView local_log_view(LogEntry* keys, unsigned int* cursor, unsigned nbins, unsigned cap, int* error,
                              unsigned long long* hpoly) {
    LogView lg{};
    lg.owner[0] = keys;
    lg.cursor = cursor; lg.nbins = nbins; lg.cap = cap; lg.lp_shift = 31; lg.src = 0; lg.error = error; lg.hpoly = hpoly;
    return lg;
}
static LogView log_view(tg_table* t) {
    return local_log_view(t->log.keys, t->log.cursor, t->log.nbins, t->log.cap, t->d_error, t->log.hpoly);
}

// replay + reset on stream 0 (stream-ordered; no host sync)
// coarse log -> one segment per partition (stream-ordered); false when the fine log cannot be had (then the coarse bins
// are replayed as they are: correct, but a bin's partitions no longer fit in L2)
static bool refine_log_async(tg_table* t) {
    tg_ctx* c = t->ctx;
    KeyLog& lg = t->log;
    if (!log_refines(t, lg.nbins)) return false;
    const unsigned np = t->g.nparts;
    uint64_t fcap = (uint64_t)((double)lg.cap * lg.nbins / np * 1.3) + 2048;
    fcap = (fcap + LOG_CAP_ALIGN - 1) / LOG_CAP_ALIGN * LOG_CAP_ALIGN;
    if (fcap > LOG_CAP_MAX) return false;
    if (lg.fine_bins != np || lg.fine_cap < fcap) {
        if (lg.fine_keys) cudaFree(lg.fine_keys);
        if (lg.fine_cursor) cudaFree(lg.fine_cursor);
        lg.fine_keys = nullptr; lg.fine_cursor = nullptr; lg.fine_bins = 0; lg.fine_cap = 0;
        if (cudaMalloc(&lg.fine_keys, (uint64_t)np * fcap * sizeof(LogEntry)) != cudaSuccess) { cudaGetLastError(); lg.fine_keys = nullptr; return false; }
        if (cudaMalloc(&lg.fine_cursor, np * sizeof(unsigned int)) != cudaSuccess) {
            cudaGetLastError(); cudaFree(lg.fine_keys); lg.fine_keys = nullptr; lg.fine_cursor = nullptr; return false;
        }
        lg.fine_bins = np; lg.fine_cap = (unsigned)fcap;
    }
    if (lg.plan_bins < np) {
        unsigned long long* p = nullptr;
        if (cudaMalloc(&p, log_replay_plan_words(1, np, 64) * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return false; }
        cudaStreamSynchronize(c->stream[0]);
        cudaFree(lg.chunk_start);
        lg.chunk_start = p; lg.plan_bins = np;
    }
    if (cudaMemsetAsync(lg.fine_cursor, 0, np * sizeof(unsigned int), c->stream[0]) != cudaSuccess) return false;
    if (launch_log_refine(lg.keys, lg.cursor, lg.cap, 1, lg.nbins, lg.chunk_start, lg.fine_keys, lg.fine_cursor, lg.fine_cap, np,
                          0, np, t->d_error, t->view(), c->sm_count, c->stream[0]) != cudaSuccess) return false;
    c->launches += 2;
    return true;
}

static int replay_log_async(tg_table* t) {
    tg_ctx* c = t->ctx;
    if (refine_log_async(t)) {
        CU(launch_log_replay(t->log.fine_keys, t->log.fine_cursor, t->log.fine_cap, 1, t->log.fine_bins, 0, t->log.fine_bins,
                             c->replay_groups, t->log.chunk_start, t->log.hpoly, t->view(),
                             c->replay_prefetch | (c->replay_fold ? 2 : 0), c->sm_count, c->stream[0]));
        c->launches += 2;
        CU(cudaMemsetAsync(t->log.cursor, 0, t->log.nbins * sizeof(unsigned int), c->stream[0]));
        t->log.pending_ub = 0;
        return TG_OK;
    }
    CU(launch_log_replay(t->log.keys, t->log.cursor, t->log.cap, 1, t->log.nbins, 0, t->log.nbins, c->replay_groups,
                         t->log.chunk_start, t->log.hpoly, t->view(), c->replay_prefetch | (c->replay_fold ? 2 : 0), c->sm_count, c->stream[0]));
    c->launches += 2;
    CU(cudaMemsetAsync(t->log.cursor, 0, t->log.nbins * sizeof(unsigned int), c->stream[0]));
    t->log.pending_ub = 0;
    return TG_OK;
}

// Distinct keys in the log, estimated from a hash-uniform sample of bins: the sample is replayed into a scratch
// table and its exact distinct count scaled up.  Keys spread over bins by hash, so the estimate is tight whatever
// the multiplicity skew; it counts keys already in the table too, i.e. it errs high.
static int estimate_log_distinct(tg_table* t, const std::vector<unsigned>& fill, uint64_t total, uint64_t* est) {
    tg_ctx* c = t->ctx;
    const unsigned nbins = t->log.nbins;
    const uint64_t mean = total / nbins + 1;
    unsigned ns = (unsigned)std::min<uint64_t>(nbins, (4000000 + mean - 1) / mean);
    if (ns < 1) ns = 1;
    uint64_t sample = 0;
    for (unsigned b = 0; b < ns; b++) sample += fill[b];
    const uint64_t per_entry = (uint64_t)le_max_run(t->k);
    if (sample == 0) { *est = total * per_entry; return TG_OK; }
    Geo sg;
    sg.subcap = whole_buckets(sample * per_entry * 2 + 1024); sg.nparts = 1; sg.part0 = 0; sg.nlocal = 1; sg.k = t->k; sg.floor = 0u;
    if (sg.subcap < MIN_SLOTS) sg.subcap = whole_buckets(MIN_SLOTS);
    CU(c->est_scratch.ensure(sg.subcap * sizeof(Slot) + 256));
    Slot* scratch = (Slot*)c->est_scratch.p;
    unsigned long long* d_n = (unsigned long long*)((char*)c->est_scratch.p + sg.subcap * sizeof(Slot));
    CU(cudaMemsetAsync(scratch, 0, sg.subcap * sizeof(Slot) + sizeof *d_n, c->stream[0]));
    TableView sv{scratch, sg, d_n, t->d_error};
    CU(launch_log_replay(t->log.keys, t->log.cursor, t->log.cap, 1, ns, 0, nbins, 1, t->log.chunk_start, nullptr, sv, 0,
                         c->sm_count, c->stream[0]));
    c->launches += 2;
    unsigned long long d = 0;
    CU(cudaMemcpyAsync(&d, d_n, sizeof d, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    *est = (uint64_t)((double)d * nbins / ns * 1.03) + 65536;
    if (*est > total * per_entry) *est = total * per_entry;
    return TG_OK;
}

// Apply every pending log entry to the table (growing it first if the log could overfill it).
// only_stream0: everything that fed the log ran on stream 0 and stream 1 is busy with something the flush must not wait
// for (the uploads of a held buffer).
static int flush_log(tg_table* t, bool only_stream0) {
    if (!t->log.keys || t->log.pending_ub == 0) return TG_OK;
    tg_ctx* c = t->ctx;
    int rc;
    if (only_stream0) CU(cudaStreamSynchronize(c->stream[0]));
    else if ((rc = sync_all(c))) return rc;
    if ((rc = table_refresh(t))) return rc;
    std::vector<unsigned> fill(t->log.nbins);
    CU(cudaMemcpy(fill.data(), t->log.cursor, fill.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    uint64_t total = 0;
    for (auto& f : fill) { if (f > t->log.cap) f = t->log.cap; total += f; }
    const uint64_t kmers_ub = total * (uint64_t)le_max_run(t->k);          // an entry holds up to that many k-mers
    if ((double)(t->distinct_ub + kmers_ub) > MAX_LOAD * (double)t->cap) {
        uint64_t est = kmers_ub;
        if ((rc = estimate_log_distinct(t, fill, total, &est))) return rc;
        if ((double)(t->distinct_ub + est) > MAX_LOAD * (double)t->cap) {
            Geo ng;
            if ((rc = grown_geo(t, t->distinct_ub + est, &ng))) return rc;
            if ((rc = table_regrow(t, ng))) return rc;
        }
    }
    if ((rc = replay_log_async(t))) return rc;
    CU(cudaStreamSynchronize(c->stream[0]));
    return table_refresh(t);
}

extern "C" {

int tg_records_hold(tg_ctx* c, const char* recs, uint64_t nbytes) {
    if (!c || (!recs && nbytes)) return fail(TG_ERR_ARG, "tg_records_hold: null argument");
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    if (rc) return rc;
    c->held.host = nullptr; c->held.nbytes = 0; c->held.uploaded = false;
    c->locus_held.forget();
    if (nbytes == 0) return TG_OK;
    CU(c->held.dev.ensure(padded_record_bytes(nbytes)));
    c->held.host = recs; c->held.nbytes = nbytes;
    return TG_OK;
}

int tg_records_release(tg_ctx* c) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    c->held.host = nullptr; c->held.nbytes = 0; c->held.uploaded = false;
    c->locus_held.forget();
    return rc;
}

int tg_records_pin_dev(tg_ctx* c, const void* d_recs, const void* d_offs, uint64_t nreads) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    c->locus_pin.forget();
    c->pin_recs = d_recs; c->pin_offs = d_offs; c->pin_nreads = d_recs && d_offs ? nreads : 0;
    return rc;
}

}  // extern "C"

static bool is_held(const tg_ctx* c, const char* recs, uint64_t nbytes) {
    return c->held.host && recs == c->held.host && nbytes == c->held.nbytes;
}

// First use of a held buffer: upload it in chunks of whole tiles on stream 1 while `consume(device pointer, bytes)` works on
// the chunks behind on stream 0.  A chunk's kernel reads a few bytes past its end (the halo of its last tile), so chunk i is
// consumed only after chunk i + 1 has arrived; the tail of the buffer is '\n' padding.  Later uses find the copy in place.
template <typename Consume>
static int held_stream(tg_ctx* c, Consume consume) {
    auto& h = c->held;
    char* dev = (char*)h.dev.p;
    const uint64_t chunk = std::max<uint64_t>(c->held_chunk_bytes, CT_TILE) / CT_TILE * CT_TILE;
    if (h.uploaded) {
        for (uint64_t pos = 0; pos < h.nbytes; pos += chunk) {
            int rc = consume(dev + pos, std::min(chunk, h.nbytes - pos));
            if (rc) return rc;
        }
        return TG_OK;
    }
    // every upload is queued on stream 1 before the first chunk is consumed: a consumer that blocks the host (a log flush
    // with its read-backs) then never holds the copies up
    const uint64_t padded = padded_record_bytes(h.nbytes);
    CU(cudaMemsetAsync(dev + h.nbytes, '\n', padded - h.nbytes, c->stream[1]));
    const uint64_t nchunks = (h.nbytes + chunk - 1) / chunk;
    while (h.chunk_events.size() < nchunks) {
        cudaEvent_t e = nullptr;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h.chunk_events.push_back(e);
    }
    for (uint64_t i = 0; i < nchunks; i++) {
        const uint64_t pos = i * chunk, n = std::min(chunk, h.nbytes - pos);
        CU(cudaMemcpyAsync(dev + pos, h.host + pos, n, cudaMemcpyHostToDevice, c->stream[1]));
        CU(cudaEventRecord(h.chunk_events[i], c->stream[1]));
    }
    for (uint64_t i = 0; i < nchunks; i++) {
        // chunk i is complete together with its halo once chunk i + 1 has arrived (or the padding behind the last one)
        CU(cudaStreamWaitEvent(c->stream[0], h.chunk_events[std::min(i + 1, nchunks - 1)], 0));
        const uint64_t pos = i * chunk;
        int rc = consume(dev + pos, std::min(chunk, h.nbytes - pos));
        if (rc) return rc;
    }
    h.uploaded = true;
    return TG_OK;
}

extern "C" {

int tg_count_reads(tg_table* t, const char* recs, uint64_t nbytes, int canonical) {
    if (!t || (!recs && nbytes)) return fail(TG_ERR_ARG, "tg_count_reads: null argument");
    if (t->kind != TG_TABLE_COUNT) return fail(TG_ERR_ARG, "tg_count_reads needs a TG_TABLE_COUNT table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc;
    // Partitioned path: append to the log batch by batch (H2D of one batch overlaps the log kernel of the other),
    // replay when the log is full and at the end.  A table still too small to be partitioned is first grown to the
    // size the input suggests -- only ever when the caller's hint was far below the input.
    bool logged = false;
    if (c->count_mode != 1 && !t->sharded() && nbytes >= (256ull << 20) && t->g.nparts < 4 && nbytes / 8 > t->cap) {
        if ((rc = tg_table_reserve(t, nbytes / 8))) return rc;
        t->distinct_ub -= nbytes / 8;       // it was a sizing hint, not an insertion
    }
    if (log_pays(t, nbytes) && (rc = ensure_log(t, log_launch_cost(c, t->log, nbytes), &logged))) return rc;
    const uint64_t room = logged ? log_room(t->log) : 0;
    if (is_held(c, recs, nbytes)) {
        // the caller promised the buffer does not change: one device copy, uploaded here (or found in place) and left for
        // the statistics call that follows.  The log is replayed in SEGMENTS (a quarter of the input each), so that the
        // table work starts while the upload is still running; only the first segment's flush talks to the host (exact
        // distinct count so far -> is the table large enough for the rest?), the later ones are stream-ordered.
        if ((rc = sync_all(c))) return rc;
        const uint64_t seg_bytes = std::max<uint64_t>(nbytes / 4, 256ull << 20);
        const uint64_t seg_cost = logged ? std::min<uint64_t>(room, log_launch_cost(c, t->log, seg_bytes)) : 0;
        uint64_t done_bytes = 0, seg_start = 0;
        bool async_ok = false;
        uint64_t distinct_at_seg_start = t->distinct_ub;
        auto flush_segment = [&]() -> int {
            int r2;
            if (async_ok) return replay_log_async(t);
            if ((r2 = flush_log(t, true))) return r2;
            // the segment [seg_start, done_bytes) added t->distinct_ub - distinct_at_seg_start keys; a later segment of the
            // same size adds at most about as many new ones (the expressed k-mers are in already).  With 1.5x head-room
            // on that, does everything fit below the load limit?  Then nobody needs to look again.
            const uint64_t seg_len = std::max<uint64_t>(done_bytes - seg_start, 1);
            const double per_byte = (double)(t->distinct_ub - distinct_at_seg_start) / (double)seg_len;
            const double projected = (double)t->distinct_ub + 1.5 * per_byte * (double)(nbytes - done_bytes);
            async_ok = projected <= MAX_LOAD * (double)t->cap;
            seg_start = done_bytes; distinct_at_seg_start = t->distinct_ub;
            return TG_OK;
        };
        rc = held_stream(c, [&](const char* d, uint64_t n) -> int {
            int r2;
            if (logged) {
                if (t->log.pending_ub && t->log.pending_ub + log_launch_cost(c, t->log, n) > seg_cost && (r2 = flush_segment())) return r2;
                CU(launch_log_tiles((const uint8_t*)d, n, t->k, canonical, log_view(t), t->view(), c->sm_count, c->stream[0]));
                t->log.pending_ub += log_launch_cost(c, t->log, n);
            } else {
                if ((r2 = tg_table_reserve(t, n))) return r2;
                CU(launch_count_tiles((const uint8_t*)d, n, t->k, canonical, t->view(), c->sm_count, c->stream[0]));
            }
            c->launches++;
            done_bytes += n;
            return TG_OK;
        });
        if (rc) return rc;
        if (logged && t->log.pending_ub && (rc = flush_segment())) return rc;
        if ((rc = sync_all(c))) return rc;
        return table_refresh(t);
    }
    uint64_t pos = 0;
    for (int it = 0; pos < nbytes; it++) {
        const int b = it & 1;
        const uint64_t end = batch_end(recs, pos, nbytes, c->batch_bytes);
        const uint64_t n = end - pos;
        if (logged) {
            if (t->log.pending_ub + log_launch_cost(c, t->log, n) > room && t->log.pending_ub) {
                if ((rc = flush_log(t))) return rc;
            }
        } else {
            if ((rc = tg_table_reserve(t, n))) return rc;     // may rehash: syncs both streams itself
        }
        CU(cudaStreamSynchronize(c->stream[b]));          // staging buffer b is free again
        if ((rc = upload_records(c, b, recs + pos, n))) return rc;
        if (logged) {
            CU(launch_log_tiles((const uint8_t*)c->recs[b].p, n, t->k, canonical, log_view(t), t->view(), c->sm_count,
                                c->stream[b]));
            t->log.pending_ub += log_launch_cost(c, t->log, n);
        } else {
            CU(launch_count_tiles((const uint8_t*)c->recs[b].p, n, t->k, canonical, t->view(), c->sm_count, c->stream[b]));
        }
        c->launches++;
        pos = end;
    }
    if (logged && (rc = flush_log(t))) return rc;
    if ((rc = sync_all(c))) return rc;
    return table_refresh(t);
}

int tg_count_reads_dev(tg_table* t, const void* d_recs, uint64_t nbytes, int canonical) {
    if (!t || !d_recs) return fail(TG_ERR_ARG, "tg_count_reads_dev: null argument");
    if (t->kind != TG_TABLE_COUNT) return fail(TG_ERR_ARG, "tg_count_reads_dev needs a TG_TABLE_COUNT table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    // the caller sizes the table (tg_table_create / tg_table_reserve): a conservative per-byte bound would
    // multiply the footprint.  An overflow raises the table's error flag -> TG_ERR_TABLE at tg_table_info.
    int rc;
    bool logged = false;
    if (log_pays(t, nbytes) && (rc = ensure_log(t, log_launch_cost(c, t->log, nbytes), &logged))) return rc;
    if (!logged) {
        CU(launch_count_tiles((const uint8_t*)d_recs, nbytes, t->k, canonical, t->view(), c->sm_count, c->stream[0]));
        c->launches++;
        return TG_OK;
    }
    // segments of whole tiles, each small enough for the log; everything stays stream-ordered on stream 0
    // one segment when the log was laid out for this input (the usual case); else equal segments of whole tiles
    const uint64_t room = std::max<uint64_t>(log_room(t->log), 1);
    const uint64_t nseg = (log_launch_cost(c, t->log, nbytes) + room - 1) / room;
    uint64_t seg = ((nbytes + nseg - 1) / nseg + CT_TILE - 1) / CT_TILE * CT_TILE;
    if (seg < (uint64_t)CT_TILE) seg = CT_TILE;
    for (uint64_t pos = 0; pos < nbytes; pos += seg) {
        const uint64_t n = std::min(seg, nbytes - pos);
        CU(launch_log_tiles((const uint8_t*)d_recs + pos, n, t->k, canonical, log_view(t), t->view(), c->sm_count,
                            c->stream[0]));
        c->launches++;
        t->log.pending_ub += log_launch_cost(c, t->log, n);
        if ((rc = replay_log_async(t))) return rc;
    }
    return TG_OK;
}

int tg_locus_prepare_dev(tg_ctx* c, int k, int recompute) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    if (!c->pin_nreads) return fail(TG_ERR_ARG, "tg_locus_prepare_dev: no device record buffer is pinned (tg_records_pin_dev)");
    // on stream 1, behind everything queued on stream 0 so far (an earlier consumer may still be reading the old order);
    // the consumers on stream 0 wait for the `ready` event, so the scan and the sort run beside whatever stream 0 does next
    if (recompute) c->locus_pin.forget();
    if (!c->order[0]) {
        CU(cudaEventCreateWithFlags(&c->order[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->order[1], cudaEventDisableTiming));
    }
    CU(cudaEventRecord(c->order[0], c->stream[0]));
    CU(cudaStreamWaitEvent(c->stream[1], c->order[0], 0));
    const uint32_t* ord = nullptr;
    return locus_order_async(c, 1, (const uint8_t*)c->pin_recs, (const uint64_t*)c->pin_offs, 0, c->pin_nreads, k, &ord);
}

int tg_count_records_dev(tg_table* t, const void* d_recs, const void* d_offs, uint64_t nreads, int canonical) {
    if (!t || !d_recs || !d_offs) return fail(TG_ERR_ARG, "tg_count_records_dev: null argument");
    if (t->kind != TG_TABLE_COUNT) return fail(TG_ERR_ARG, "tg_count_records_dev needs a TG_TABLE_COUNT table");
    if (t->sharded()) return fail(TG_ERR_ARG, "tg_count_records_dev: sharded tables are counted through the exchange path");
    if (nreads > 0x7FFFFFF0ull) return fail(TG_ERR_ARG, "tg_count_records_dev: at most 2^31 reads per call");
    if (t->k > 31) return fail(TG_ERR_ARG, "tg_count_records_dev: k = 32 is counted by tg_count_reads / tg_count_reads_dev");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (t->log.pending_ub) { int rc = flush_log(t); if (rc) return rc; }
    // stream-ordered like tg_count_reads_dev; the caller sizes the table
    const uint32_t* ord = nullptr;
    if (int rc = locus_order_async(c, 0, (const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, t->k, &ord)) return rc;
    CU(launch_count_reads((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, t->k, canonical, t->view(), ord, c->stream[0]));
    c->launches++;
    return TG_OK;
}

// ---- sharded counting (multi-GPU): phase 1 into a caller-owned log, phase 2 from a caller-owned (received) log ----
int tg_count_partition_dev(tg_ctx* c, const void* d_recs, uint64_t nbytes, int k, int canonical, uint32_t nbins,
                           uint32_t cap, void* d_keys, void* d_cursor, void* d_hpoly) {
    if (!c || !d_recs || !d_keys || !d_cursor || !d_hpoly)
        return fail(TG_ERR_ARG, "tg_count_partition_dev: null argument");
    if (k < MIN_FAST_K || k > 31) return fail(TG_ERR_ARG, "k-mer length %d unsupported by the partitioned path (%d..31)", k, MIN_FAST_K);
    if (nbins == 0 || nbins > LOG_MAX_BINS || cap == 0 || cap > LOG_CAP_MAX)
        return fail(TG_ERR_ARG, "tg_count_partition_dev: bad log shape (1..%u bins, capacity at most %u)", LOG_MAX_BINS,
                    LOG_CAP_MAX);
    if (bind(c)) return TG_ERR_CUDA;
    LogView lg = local_log_view((LogEntry*)d_keys, (unsigned int*)d_cursor, nbins, cap, c->d_error,
                                (unsigned long long*)d_hpoly);
    TableView none{nullptr, Geo{0, 1, 0, 1, k, 0u}, nullptr, nullptr};
    CU(launch_log_tiles((const uint8_t*)d_recs, nbytes, k, canonical, lg, none, c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}

// fused phase 1 + exchange: entries go straight into the owners' receive logs through peer memory
int tg_count_partition_peers_dev(tg_ctx* c, const void* d_recs, uint64_t nbytes, int k, int canonical, uint32_t nbins,
                                 uint32_t cap, uint32_t nranks, uint32_t my_rank, void* const* d_owner_keys,
                                 void* d_cursor, void* d_hpoly) {
    if (!c || !d_recs || !d_owner_keys || !d_cursor || !d_hpoly)
        return fail(TG_ERR_ARG, "tg_count_partition_peers_dev: null argument");
    if (k < MIN_FAST_K || k > 31) return fail(TG_ERR_ARG, "k-mer length %d unsupported by the partitioned path (%d..31)", k, MIN_FAST_K);
    if (nranks == 0 || nranks > (uint32_t)LOG_MAX_RANKS || my_rank >= nranks)
        return fail(TG_ERR_ARG, "tg_count_partition_peers_dev: 1..%d ranks, rank inside", LOG_MAX_RANKS);
    if (nbins == 0 || nbins > LOG_MAX_BINS || nbins % nranks || cap == 0 || cap > LOG_CAP_MAX)
        return fail(TG_ERR_ARG, "tg_count_partition_peers_dev: bad log shape (bins a multiple of the ranks, at most %u)",
                    LOG_MAX_BINS);
    const uint32_t lp = nbins / nranks;
    if (lp & (lp - 1)) return fail(TG_ERR_ARG, "tg_count_partition_peers_dev: bins per rank must be a power of two");
    if (bind(c)) return TG_ERR_CUDA;
    LogView lg{};
    for (uint32_t r = 0; r < nranks; r++) {
        if (!d_owner_keys[r]) return fail(TG_ERR_ARG, "tg_count_partition_peers_dev: null receive log for rank %u", r);
        lg.owner[r] = (LogEntry*)d_owner_keys[r];
    }
    unsigned sh = 0;
    while ((1u << sh) < lp) sh++;
    lg.cursor = (unsigned int*)d_cursor; lg.nbins = nbins; lg.cap = cap; lg.lp_shift = sh; lg.src = my_rank;
    lg.error = c->d_error; lg.hpoly = (unsigned long long*)d_hpoly;
    TableView none{nullptr, Geo{0, 1, 0, 1, k, 0u}, nullptr, nullptr};
    CU(launch_log_tiles((const uint8_t*)d_recs, nbytes, k, canonical, lg, none, c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}

// owner-side split of a received coarse log into the table's partitions (see k_log_refine)
// ---- routed lookups (multi-GPU statistics against a sharded table that is NOT replicated) -------------------------------
// requester: keys binned by owner like counted k-mers (same bins, same log layout), return addresses kept locally
int tg_query_partition_dev(tg_ctx* c, const void* d_recs, uint64_t nbytes, int k, int canonical, uint32_t nbins, uint32_t cap,
                           void* d_keys, void* d_cursor, void* d_posidx) {
    if (!c || !d_recs || !d_keys || !d_cursor || !d_posidx) return fail(TG_ERR_ARG, "tg_query_partition_dev: null argument");
    if (k < 1 || k > 31) return fail(TG_ERR_ARG, "k-mer length %d unsupported (1..31)", k);
    if (nbins == 0 || nbins > LOG_MAX_BINS || cap == 0 || cap > LOG_CAP_MAX)
        return fail(TG_ERR_ARG, "tg_query_partition_dev: bad log shape (1..%u bins, capacity at most %u)", LOG_MAX_BINS, LOG_CAP_MAX);
    if (nbytes > 0xFFFF0000ull) return fail(TG_ERR_ARG, "tg_query_partition_dev: at most 4 GiB of records per call");
    if (bind(c)) return TG_ERR_CUDA;
    LogView lg = local_log_view((LogEntry*)d_keys, (unsigned int*)d_cursor, nbins, cap, c->d_error, nullptr);
    lg.posidx = (unsigned int*)d_posidx;
    TableView none{nullptr, Geo{0, 1, 0, 1, k, 0u}, nullptr, nullptr};
    CU(launch_log_tiles((const uint8_t*)d_recs, nbytes, k, canonical, lg, none, c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}
// owner: received keys [nsrc][lp][cap] (lp = bins this rank owns) -> d_resp, same layout, the table's value or 0
int tg_query_answer_dev(tg_table* t, const void* d_keys, const void* d_cursor, uint32_t nsrc, uint32_t lp, uint32_t cap,
                        void* d_resp) {
    if (!t || !d_keys || !d_cursor || !d_resp) return fail(TG_ERR_ARG, "tg_query_answer_dev: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (t->log.pending_ub) { int rc = flush_log(t); if (rc) return rc; }
    CU(launch_query_answer((const unsigned long long*)d_keys, (const unsigned int*)d_cursor, cap, nsrc, lp, t->slots, t->g,
                           (unsigned int*)d_resp, c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}
// requester: answers (its own log layout [nbins][cap]) -> d_counts[position of the window in the record buffer]
int tg_query_scatter_dev(tg_ctx* c, const void* d_resp, const void* d_posidx, const void* d_cursor, uint32_t nbins, uint32_t cap,
                         void* d_counts) {
    if (!c || !d_resp || !d_posidx || !d_cursor || !d_counts) return fail(TG_ERR_ARG, "tg_query_scatter_dev: null argument");
    if (bind(c)) return TG_ERR_CUDA;
    CU(launch_query_scatter((const unsigned int*)d_resp, (const unsigned int*)d_posidx, (const unsigned int*)d_cursor, nbins, cap,
                            (unsigned int*)d_counts, c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}
// median / mean / stdev of every read from counts that sit at the windows' positions (u32 per byte of the record buffer,
// 0 = absent or never asked); count_floor as in tg_table_set_count_floor
int tg_cov_stats_counts_dev(tg_ctx* c, const void* d_recs, const void* d_offs, uint64_t nreads, int k, uint32_t count_floor,
                            const void* d_counts, void* d_median, void* d_mean, void* d_stdev) {
    if (!c || !d_recs || !d_offs || !d_counts || !d_median || !d_mean || !d_stdev)
        return fail(TG_ERR_ARG, "tg_cov_stats_counts_dev: null argument");
    if (k < 1 || k > 31) return fail(TG_ERR_ARG, "k-mer length %d unsupported (1..31)", k);
    if (nreads > 0x7FFFFFF0ull) return fail(TG_ERR_ARG, "tg_cov_stats_counts_dev: at most 2^31 reads per call");
    if (bind(c)) return TG_ERR_CUDA;
    const int b = 0;
    CU(c->long_idx[b].ensure(nreads * 4));
    CU(c->long_scratch.ensure(c->long_scratch_bytes));
    CU(cudaMemsetAsync(c->d_long_hdr[b], 0, 2 * sizeof(unsigned int), c->stream[b]));
    LongList ll{c->d_long_hdr[b], c->d_long_hdr[b] + 1, (unsigned int*)c->long_idx[b].p};
    Geo g{0, 1, 0, 1, k, count_floor};
    CU(launch_cov_stats((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, k, 1, nullptr, g, (uint32_t*)d_median,
                        (float*)d_mean, (float*)d_stdev, nullptr, ll, nullptr, c->stats_arena, (const uint32_t*)d_counts, c->stream[b]));
    CU(launch_cov_stats_long_auto((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, k, 1, nullptr, g, (uint32_t*)d_median,
                                  (float*)d_mean, (float*)d_stdev, nullptr, ll, c->long_scratch.p, c->long_scratch_bytes, c->d_error,
                                  c->sm_count * 2, c->stream[b], (const uint32_t*)d_counts));
    c->launches += 2;
    return TG_OK;
}

int tg_log_refine_dev(tg_ctx* c, const void* d_keys, const void* d_cursor, uint32_t nsrc, uint32_t ncoarse, uint32_t cap,
                      void* d_out_keys, void* d_out_cursor, uint32_t nfine, uint32_t out_cap, uint32_t fine0,
                      uint32_t nfine_global) {
    if (!c || !d_keys || !d_cursor || !d_out_keys || !d_out_cursor)
        return fail(TG_ERR_ARG, "tg_log_refine_dev: null argument");
    if (nsrc == 0 || ncoarse == 0 || cap == 0 || nfine == 0 || out_cap == 0 || out_cap > LOG_CAP_MAX || nfine % ncoarse ||
        (uint64_t)fine0 + nfine > nfine_global || nfine / ncoarse > log_refine_max_split() || nfine > LOG_MAX_BINS)
        return fail(TG_ERR_ARG, "tg_log_refine_dev: bad shape (fine bins a multiple of the coarse bins, at most %u per "
                                "coarse bin, inside the global partition range)", log_refine_max_split());
    if (bind(c)) return TG_ERR_CUDA;
    const size_t need = std::max(log_refine_plan_words(nsrc, ncoarse), log_replay_plan_words(1, nfine, 64)) *
                        sizeof(unsigned long long);
    CU(c->scratch.ensure(need));
    CU(launch_log_refine((const LogEntry*)d_keys, (const unsigned int*)d_cursor, cap, nsrc, ncoarse,
                         (unsigned long long*)c->scratch.p, (LogEntry*)d_out_keys, (unsigned int*)d_out_cursor,
                         out_cap, nfine, fine0, nfine_global, c->d_error, TableView{nullptr, Geo{0, 1, 0, 1, 0, 0u}, nullptr, nullptr},
                         c->sm_count, c->stream[0]));
    c->launches += 2;
    return TG_OK;
}

// ---- peer memory: CUDA IPC handles of allocations made with tg_dev_alloc --------------------------------------------
int tg_ipc_export(tg_ctx* c, void* dptr, uint8_t* handle) {
    if (!c || !dptr || !handle) return fail(TG_ERR_ARG, "tg_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == TG_IPC_HANDLE_BYTES, "IPC handle size");
    if (bind(c)) return TG_ERR_CUDA;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, sizeof h);
    return TG_OK;
}

int tg_ipc_open(tg_ctx* c, const uint8_t* handle, void** dptr) {
    if (!c || !handle || !dptr) return fail(TG_ERR_ARG, "tg_ipc_open: null argument");
    if (bind(c)) return TG_ERR_CUDA;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TG_OK;
}

int tg_ipc_close(tg_ctx* c, void* dptr) {
    if (!c || !dptr) return fail(TG_ERR_ARG, "tg_ipc_close: null argument");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaIpcCloseMemHandle(dptr));
    return TG_OK;
}

int tg_table_replay_log_dev(tg_table* t, const void* d_keys, const void* d_cursor, void* d_hpoly, uint32_t nsrc,
                            uint32_t cap) {
    if (!t || !d_keys || !d_cursor || nsrc == 0 || cap == 0) return fail(TG_ERR_ARG, "tg_table_replay_log_dev: bad argument");
    if (t->kind != TG_TABLE_COUNT) return fail(TG_ERR_ARG, "tg_table_replay_log_dev needs a TG_TABLE_COUNT table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    const size_t need = log_replay_plan_words(nsrc, t->g.nlocal, 64) * sizeof(unsigned long long);
    CU(c->scratch.ensure(need));
    CU(launch_log_replay((const LogEntry*)d_keys, (const unsigned int*)d_cursor, cap, nsrc, t->g.nlocal, t->g.part0,
                         t->g.nparts, c->replay_groups, (unsigned long long*)c->scratch.p, (unsigned long long*)d_hpoly,
                         t->view(), c->replay_prefetch | (c->replay_fold ? 2 : 0), c->sm_count, c->stream[0]));
    c->launches += 2;
    return TG_OK;
}

int tg_table_load_pairs(tg_table* t, const uint64_t* keys, const uint32_t* vals, uint64_t n, int canonical) {
    if (!t || ((!keys || !vals) && n)) return fail(TG_ERR_ARG, "tg_table_load_pairs: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc;
    if ((rc = flush_log(t))) return rc;
    const uint64_t chunk = std::max<uint64_t>(1, c->batch_bytes / 12);
    for (uint64_t pos = 0, it = 0; pos < n; pos += chunk, it++) {
        const int b = (int)(it & 1);
        const uint64_t m = std::min(chunk, n - pos);
        if ((rc = tg_table_reserve(t, m))) return rc;
        CU(cudaStreamSynchronize(c->stream[b]));
        CU(c->out_a[b].ensure(m * 8));
        CU(c->out_b[b].ensure(m * 4));
        CU(cudaMemcpyAsync(c->out_a[b].p, keys + pos, m * 8, cudaMemcpyHostToDevice, c->stream[b]));
        CU(cudaMemcpyAsync(c->out_b[b].p, vals + pos, m * 4, cudaMemcpyHostToDevice, c->stream[b]));
        CU(launch_load_pairs((const uint64_t*)c->out_a[b].p, (const uint32_t*)c->out_b[b].p, m, t->k, canonical,
                             t->view(), t->kind == TG_TABLE_LABEL, c->stream[b]));
        c->launches++;
    }
    if ((rc = sync_all(c))) return rc;
    return table_refresh(t);
}

int tg_table_export(tg_table* t, uint32_t min_count, uint32_t max_count, int sorted, int canonical_repr,
                    uint64_t** keys, uint32_t** counts, uint64_t* n_out) {
    if (!t || !keys || !counts || !n_out) return fail(TG_ERR_ARG, "tg_table_export: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    *keys = nullptr; *counts = nullptr; *n_out = 0;
    int rc;
    if ((rc = flush_log(t))) return rc;
    if ((rc = sync_all(c))) return rc;
    if ((rc = table_refresh(t))) return rc;
    cudaStream_t s = c->stream[0];
    unsigned long long* d_n = nullptr;
    CU(cudaMalloc(&d_n, sizeof *d_n));
    CU(cudaMemsetAsync(d_n, 0, sizeof *d_n, s));
    CU(launch_export(t->slots, t->cap, min_count, max_count, t->k, canonical_repr, nullptr, nullptr, d_n, s));
    c->launches++;
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (n == 0) { cudaFree(d_n); *keys = (uint64_t*)malloc(8); *counts = (uint32_t*)malloc(4); return TG_OK; }
    uint64_t* d_keys = nullptr; uint32_t* d_vals = nullptr;
    CU(cudaMalloc(&d_keys, n * 8));
    CU(cudaMalloc(&d_vals, n * 4));
    CU(cudaMemsetAsync(d_n, 0, sizeof *d_n, s));
    CU(launch_export(t->slots, t->cap, min_count, max_count, t->k, canonical_repr, d_keys, d_vals, d_n, s));
    c->launches++;
    if (sorted) { CU(sort_pairs(d_keys, d_vals, n, t->k, s)); c->launches += 8; }
    uint64_t* hk = (uint64_t*)malloc(n * 8);
    uint32_t* hv = (uint32_t*)malloc(n * 4);
    if (!hk || !hv) { free(hk); free(hv); cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_n);
                      return fail(TG_ERR_NOMEM, "tg_table_export: host allocation of %llu pairs failed", n); }
    CU(cudaMemcpyAsync(hk, d_keys, n * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(hv, d_vals, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    cudaFree(d_keys); cudaFree(d_vals); cudaFree(d_n);
    *keys = hk; *counts = hv; *n_out = n;
    return TG_OK;
}

int tg_table_count_sum(tg_table* t, uint64_t* sum) {
    if (!t || !sum) return fail(TG_ERR_ARG, "tg_table_count_sum: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc;
    if ((rc = flush_log(t))) return rc;
    if ((rc = sync_all(c))) return rc;
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof *d));
    CU(cudaMemsetAsync(d, 0, sizeof *d, c->stream[0]));
    CU(launch_table_sum(t->slots, t->cap, d, c->stream[0]));
    c->launches++;
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, d, sizeof v, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    cudaFree(d);
    *sum = v;
    return TG_OK;
}

int tg_valid_windows_dev(tg_ctx* c, const void* d_recs, uint64_t nbytes, int k, uint64_t* n) {
    if (!c || !d_recs || !n || k < 1 || k > 32) return fail(TG_ERR_ARG, "tg_valid_windows_dev: bad argument");
    if (bind(c)) return TG_ERR_CUDA;
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof *d));
    CU(cudaMemsetAsync(d, 0, sizeof *d, c->stream[0]));
    CU(launch_valid_windows((const uint8_t*)d_recs, nbytes, k, d, c->stream[0]));
    c->launches++;
    unsigned long long v = 0;
    CU(cudaMemcpyAsync(&v, d, sizeof v, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    cudaFree(d);
    *n = v;
    return TG_OK;
}

int tg_histo(tg_table* t, uint64_t bins[TG_HISTO_BINS]) {
    if (!t || !bins) return fail(TG_ERR_ARG, "tg_histo: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc;
    if ((rc = flush_log(t))) return rc;
    if ((rc = sync_all(c))) return rc;
    unsigned long long* d_bins = nullptr;
    CU(cudaMalloc(&d_bins, TG_HISTO_BINS * sizeof *d_bins));
    CU(cudaMemsetAsync(d_bins, 0, TG_HISTO_BINS * sizeof *d_bins, c->stream[0]));
    CU(launch_histo(t->slots, t->cap, d_bins, c->stream[0]));
    c->launches++;
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    CU(cudaMemcpyAsync(bins, d_bins, TG_HISTO_BINS * sizeof *d_bins, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    cudaFree(d_bins);
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stage S / stage R: batches of whole records, double buffered over the two streams
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"

// One per-read pass over a held record buffer: offsets up, kernels on the device copy, three 4-byte results per read down.
// *done = false (and nothing written) when a read was too long for the fixed scratch of the device-driven long kernels.
template <typename Launch, typename A, typename B, typename C3>
static int held_pass(tg_ctx* c, const uint64_t* offs, uint64_t nreads, bool* done, Launch launch, A* out_a, B* out_b, C3* out_c) {
    auto& h = c->held;
    *done = false;
    int rc;
    if (!h.uploaded && (rc = held_stream(c, [](const char*, uint64_t) -> int { return TG_OK; }))) return rc;
    CU(h.offs.ensure((nreads + 1) * 8));
    CU(h.out_a.ensure(nreads * 4)); CU(h.out_b.ensure(nreads * 4)); CU(h.out_c.ensure(nreads * 4));
    CU(h.long_idx.ensure(nreads * 4));
    CU(c->long_scratch.ensure(c->long_scratch_bytes));
    CU(cudaMemcpyAsync(h.offs.p, offs, (nreads + 1) * 8, cudaMemcpyHostToDevice, c->stream[0]));
    CU(cudaMemsetAsync(c->d_long_hdr[0], 0, 2 * sizeof(unsigned int), c->stream[0]));
    LongList ll{c->d_long_hdr[0], c->d_long_hdr[0] + 1, (unsigned int*)h.long_idx.p};
    if ((rc = launch((const uint8_t*)h.dev.p, (const uint64_t*)h.offs.p, ll))) return rc;
    int err = 0;
    CU(cudaMemcpyAsync(&err, c->d_error, sizeof err, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    if (err == 4) { CU(cudaMemset(c->d_error, 0, sizeof(int))); return TG_OK; }
    CU(cudaMemcpyAsync(out_a, h.out_a.p, nreads * 4, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaMemcpyAsync(out_b, h.out_b.p, nreads * 4, cudaMemcpyDeviceToHost, c->stream[0]));
    if (out_c) CU(cudaMemcpyAsync(out_c, h.out_c.p, nreads * 4, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    *done = true;
    return TG_OK;
}

// Locus order of the reads [0, nreads) of a device record buffer, queued on stream b: *d_order = u32[nreads], or nullptr
// when the launch is too small to pay for it (or the knob is off).  Stream-ordered, no host synchronisation.  The order
// of the held buffer's device copy (offs_key = the caller's host offsets) and of a pinned device buffer is kept.
static int locus_order_async(tg_ctx* c, int b, const uint8_t* d_recs, const uint64_t* d_offs, uint64_t rec_base, uint64_t nreads,
                             int k, const uint32_t** d_order, const void* held_offs_key) {
    *d_order = nullptr;
    if (!c->locus_order || nreads < c->locus_min_reads || nreads > 0x7FFFFFF0ull) return TG_OK;
    const int m = k < c->locus_m ? k : c->locus_m;
    tg_ctx::LocusCache* lc = nullptr;
    const void* offs_key = d_offs;
    if (held_offs_key && c->held.uploaded && d_recs == (const uint8_t*)c->held.dev.p) { lc = &c->locus_held; offs_key = held_offs_key; }
    else if (rec_base == 0 && c->pin_nreads && d_recs == c->pin_recs && d_offs == c->pin_offs && nreads == c->pin_nreads) lc = &c->locus_pin;
    if (lc && lc->valid && lc->recs == d_recs && lc->offs_key == offs_key && lc->nreads == nreads && lc->m == m) {
        CU(cudaStreamWaitEvent(c->stream[b], lc->ready, 0));
        *d_order = lc->order;
        return TG_OK;
    }
    const size_t need = locus_sort_bytes(nreads);
    DevBuf& buf = lc ? lc->buf : c->locus[b];
    CU(buf.ensure(need));
    uint32_t* sig = (uint32_t*)buf.p;
    CU(launch_read_locus(d_recs, d_offs, rec_base, nreads, m, sig, sig + nreads, c->sm_count, c->stream[b]));
    CU(locus_sort(buf.p, need, nreads, d_order, c->stream[b]));
    c->launches += 2;
    if (lc) {
        if (!lc->ready) CU(cudaEventCreateWithFlags(&lc->ready, cudaEventDisableTiming));
        CU(cudaEventRecord(lc->ready, c->stream[b]));
        lc->recs = d_recs; lc->offs_key = offs_key; lc->nreads = nreads; lc->m = m; lc->order = *d_order; lc->valid = true;
    }
    return TG_OK;
}

struct ReadBatch { uint64_t r0, r1; };

static std::vector<ReadBatch> split_reads(const uint64_t* offs, uint64_t nreads, size_t batch_bytes) {
    std::vector<ReadBatch> v;
    uint64_t r0 = 0;
    while (r0 < nreads) {
        // largest r1 with offs[r1] - offs[r0] <= batch_bytes (at least one read)
        const uint64_t* e = std::upper_bound(offs + r0 + 1, offs + nreads + 1, offs[r0] + batch_bytes);
        uint64_t r1 = (uint64_t)(e - offs) - 1;
        if (r1 <= r0) r1 = r0 + 1;
        if (r1 - r0 > 0x7FFFFFF0ull) r1 = r0 + 0x7FFFFFF0ull;
        v.push_back({r0, r1});
        r0 = r1;
    }
    return v;
}

// largest window count among reads [r0, r1) (at least 1, so that scratch sizes are never zero)
static unsigned batch_max_windows(const uint64_t* offs, uint64_t r0, uint64_t r1, int k) {
    uint64_t mx = 1;
    for (uint64_t r = r0; r < r1; r++) {
        const uint64_t L = offs[r + 1] - offs[r] - 1;
        if (L >= (uint64_t)k && L - k + 1 > mx) mx = L - k + 1;
    }
    return (unsigned)std::min<uint64_t>(mx, 0xFFFFFFF0ull);
}
template <typename ScratchBytes>
static void long_launch_shape(tg_ctx* c, unsigned n_long, unsigned max_win, int k, ScratchBytes scratch_bytes, int* nctas,
                              size_t* need) {
    *nctas = (int)std::min<unsigned>(std::max(n_long, 1u), (unsigned)c->sm_count * 2);
    *need = scratch_bytes(max_win, k, *nctas);
    while (*nctas > 1 && *need > (1ull << 30)) { *nctas = (*nctas + 1) / 2; *need = scratch_bytes(max_win, k, *nctas); }
}

// after the warp-path kernel of batch b: run the CTA-per-read kernel if any read was too long for it
template <typename LaunchLong, typename ScratchBytes>
static int finish_long(tg_ctx* c, int b, int k, ScratchBytes scratch_bytes, LaunchLong launch_long) {
    unsigned int* h = c->h_long_hdr + 2 * b;
    CU(cudaMemcpyAsync(h, c->d_long_hdr[b], 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream[b]));
    CU(cudaStreamSynchronize(c->stream[b]));
    if (h[0] == 0) return TG_OK;
    int nctas = (int)std::min<unsigned>(h[0], (unsigned)c->sm_count * 2);
    size_t need = scratch_bytes(h[1], k, nctas);
    while (nctas > 1 && need > (1ull << 30)) { nctas = (nctas + 1) / 2; need = scratch_bytes(h[1], k, nctas); }
    // the scratch buffer is shared by both streams: make sure the other stream's long kernel is done
    CU(cudaStreamSynchronize(c->stream[b ^ 1]));
    CU(c->scratch.ensure(need));
    CU(launch_long(h[0], h[1], c->scratch.p, nctas));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream[b]));
    return TG_OK;
}

extern "C" {

int tg_cov_stats(tg_table* t, const char* recs, const uint64_t* offs, uint64_t nreads, int canonical,
                 uint32_t* median, float* mean, float* stdev, uint32_t* per_kmer) {
    if (!t || ((!recs || !offs || !median || !mean || !stdev) && nreads))
        return fail(TG_ERR_ARG, "tg_cov_stats: null argument");
    if (t->kind != TG_TABLE_COUNT) return fail(TG_ERR_ARG, "tg_cov_stats needs a TG_TABLE_COUNT table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (nreads == 0) return TG_OK;
    int rc;
    if ((rc = flush_log(t))) return rc;
    if ((rc = sync_all(c))) return rc;
    if (!per_kmer && t->k >= MIN_FAST_K && nreads <= 0x7FFFFFF0ull && is_held(c, recs, offs[nreads])) {
        // held records: the device copy is there already (or goes up once, now); one pass over all reads, no staging
        bool done = false;
        if ((rc = held_pass(c, offs, nreads, &done, [&](const uint8_t* d, const uint64_t* d_offs, LongList ll) -> int {
                const uint32_t* ord = nullptr;
                if (int r3 = locus_order_async(c, 0, d, d_offs, 0, nreads, t->k, &ord, offs)) return r3;
                CU(launch_cov_stats(d, d_offs, 0, nreads, t->k, canonical, t->slots, t->g, (uint32_t*)c->held.out_a.p,
                                    (float*)c->held.out_b.p, (float*)c->held.out_c.p, nullptr, ll, ord, c->stats_arena, nullptr, c->stream[0]));
                CU(launch_cov_stats_long_auto(d, d_offs, 0, t->k, canonical, t->slots, t->g, (uint32_t*)c->held.out_a.p,
                                              (float*)c->held.out_b.p, (float*)c->held.out_c.p, nullptr, ll, c->long_scratch.p,
                                              c->long_scratch_bytes, c->d_error, c->sm_count * 2, c->stream[0]));
                c->launches += 2;
                return TG_OK;
            }, median, mean, stdev))) return rc;
        if (done) return TG_OK;                    // else: a read too long for the fixed scratch -- the batched path below
    }
    const std::vector<ReadBatch> batches = split_reads(offs, nreads, c->batch_bytes);
    struct Pending { bool live = false; ReadBatch rb; } pend[2];
    auto drain = [&](int b) -> int {   // long-read pass + results back to the caller for the batch in flight on b
        if (!pend[b].live) return TG_OK;
        const ReadBatch rb = pend[b].rb;
        const uint64_t m = rb.r1 - rb.r0, base = offs[rb.r0];
        int r2 = finish_long(c, b, t->k, cov_stats_long_scratch_bytes,
            [&](unsigned n_long, unsigned max_win, void* scratch, int nctas) {
                return launch_cov_stats_long((const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, t->k,
                                             canonical, t->slots, t->g, (uint32_t*)c->out_a[b].p, (float*)c->out_b[b].p,
                                             (float*)c->out_c[b].p, per_kmer ? (uint32_t*)c->per_kmer[b].p : nullptr,
                                             (const unsigned int*)c->long_idx[b].p, n_long, max_win, scratch, nctas,
                                             c->stream[b]);
            });
        if (r2) return r2;
        CU(cudaMemcpyAsync(median + rb.r0, c->out_a[b].p, m * 4, cudaMemcpyDeviceToHost, c->stream[b]));
        CU(cudaMemcpyAsync(mean + rb.r0, c->out_b[b].p, m * 4, cudaMemcpyDeviceToHost, c->stream[b]));
        CU(cudaMemcpyAsync(stdev + rb.r0, c->out_c[b].p, m * 4, cudaMemcpyDeviceToHost, c->stream[b]));
        if (per_kmer)
            CU(cudaMemcpyAsync(per_kmer + base, c->per_kmer[b].p, (offs[rb.r1] - base) * 4, cudaMemcpyDeviceToHost,
                               c->stream[b]));
        CU(cudaStreamSynchronize(c->stream[b]));
        pend[b].live = false;
        return TG_OK;
    };
    for (size_t i = 0; i < batches.size(); i++) {
        const int b = (int)(i & 1);
        if ((rc = drain(b))) return rc;
        const ReadBatch rb = batches[i];
        const uint64_t m = rb.r1 - rb.r0, base = offs[rb.r0], nb = offs[rb.r1] - base;
        if ((rc = upload_records(c, b, recs + base, nb))) return rc;
        CU(c->offs[b].ensure((m + 1) * 8));
        CU(c->out_a[b].ensure(m * 4)); CU(c->out_b[b].ensure(m * 4)); CU(c->out_c[b].ensure(m * 4));
        CU(c->long_idx[b].ensure(m * 4));
        if (per_kmer) { CU(c->per_kmer[b].ensure(nb * 4)); CU(cudaMemsetAsync(c->per_kmer[b].p, 0, nb * 4, c->stream[b])); }
        CU(cudaMemcpyAsync(c->offs[b].p, offs + rb.r0, (m + 1) * 8, cudaMemcpyHostToDevice, c->stream[b]));
        CU(cudaMemsetAsync(c->d_long_hdr[b], 0, 2 * sizeof(unsigned int), c->stream[b]));
        LongList ll{c->d_long_hdr[b], c->d_long_hdr[b] + 1, (unsigned int*)c->long_idx[b].p};
        if (t->k >= MIN_FAST_K) {
            const uint32_t* ord = nullptr;
            if ((rc = locus_order_async(c, b, (const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, m, t->k, &ord))) return rc;
            CU(launch_cov_stats((const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, m, t->k, canonical,
                                t->slots, t->g, (uint32_t*)c->out_a[b].p, (float*)c->out_b[b].p, (float*)c->out_c[b].p,
                                per_kmer ? (uint32_t*)c->per_kmer[b].p : nullptr, ll, ord, c->stats_arena, nullptr, c->stream[b]));
        } else {
            // k-mers shorter than the warp path's 8 m-mers per k-mer: every read through the CTA-per-read kernel
            int nctas = 0; size_t need = 0;
            const unsigned max_win = batch_max_windows(offs, rb.r0, rb.r1, t->k);
            long_launch_shape(c, (unsigned)m, max_win, t->k, cov_stats_long_scratch_bytes, &nctas, &need);
            CU(cudaStreamSynchronize(c->stream[b ^ 1]));      // the scratch buffer is shared by both streams
            CU(c->scratch.ensure(need));
            CU(launch_cov_stats_long((const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, t->k, canonical,
                                     t->slots, t->g, (uint32_t*)c->out_a[b].p, (float*)c->out_b[b].p, (float*)c->out_c[b].p,
                                     per_kmer ? (uint32_t*)c->per_kmer[b].p : nullptr, nullptr, (unsigned)m, max_win,
                                     c->scratch.p, nctas, c->stream[b]));
        }
        c->launches++;
        pend[b].live = true; pend[b].rb = rb;
    }
    if ((rc = drain(0))) return rc;
    if ((rc = drain(1))) return rc;
    return TG_OK;
}

int tg_cov_stats_dev(tg_table* t, const void* d_recs, const void* d_offs, uint64_t nreads, int canonical,
                     void* d_median, void* d_mean, void* d_stdev) {
    if (!t || !d_recs || !d_offs || !d_median || !d_mean || !d_stdev)
        return fail(TG_ERR_ARG, "tg_cov_stats_dev: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (nreads > 0x7FFFFFF0ull) return fail(TG_ERR_ARG, "tg_cov_stats_dev: at most 2^31 reads per call");
    if (t->k < MIN_FAST_K) return fail(TG_ERR_ARG, "tg_cov_stats_dev: k >= %d (shorter k-mers: the host-buffer entry point)", MIN_FAST_K);
    if (t->log.pending_ub) { int rc = flush_log(t); if (rc) return rc; }
    // everything below is stream-ordered: no host synchronisation (see launch_cov_stats_long_auto)
    const int b = 0;
    CU(c->long_idx[b].ensure(nreads * 4));
    CU(c->long_scratch.ensure(c->long_scratch_bytes));
    CU(cudaMemsetAsync(c->d_long_hdr[b], 0, 2 * sizeof(unsigned int), c->stream[b]));
    LongList ll{c->d_long_hdr[b], c->d_long_hdr[b] + 1, (unsigned int*)c->long_idx[b].p};
    const uint32_t* ord = nullptr;
    if (int rc = locus_order_async(c, b, (const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, t->k, &ord)) return rc;
    CU(launch_cov_stats((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, t->k, canonical, t->slots, t->g,
                        (uint32_t*)d_median, (float*)d_mean, (float*)d_stdev, nullptr, ll, ord, c->stats_arena, nullptr, c->stream[b]));
    CU(launch_cov_stats_long_auto((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, t->k, canonical, t->slots, t->g,
                                  (uint32_t*)d_median, (float*)d_mean, (float*)d_stdev, nullptr, ll, c->long_scratch.p,
                                  c->long_scratch_bytes, c->d_error, c->sm_count * 2, c->stream[b]));
    c->launches += 2;
    return TG_OK;
}

int tg_label_bundles(tg_table* t, const char* recs, const uint64_t* offs, uint64_t nbundles, uint32_t first_index) {
    if (!t || ((!recs || !offs) && nbundles)) return fail(TG_ERR_ARG, "tg_label_bundles: null argument");
    if (t->kind != TG_TABLE_LABEL) return fail(TG_ERR_ARG, "tg_label_bundles needs a TG_TABLE_LABEL table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (nbundles == 0) return TG_OK;
    int rc;
    const std::vector<ReadBatch> batches = split_reads(offs, nbundles, c->batch_bytes);
    for (size_t i = 0; i < batches.size(); i++) {
        const int b = (int)(i & 1);
        const ReadBatch rb = batches[i];
        const uint64_t m = rb.r1 - rb.r0, base = offs[rb.r0], nb = offs[rb.r1] - base;
        if ((rc = tg_table_reserve(t, nb))) return rc;
        CU(cudaStreamSynchronize(c->stream[b]));
        if ((rc = upload_records(c, b, recs + base, nb))) return rc;
        CU(c->offs[b].ensure((m + 1) * 8));
        CU(cudaMemcpyAsync(c->offs[b].p, offs + rb.r0, (m + 1) * 8, cudaMemcpyHostToDevice, c->stream[b]));
        CU(launch_label_tiles((const uint8_t*)c->recs[b].p, nb, (const uint64_t*)c->offs[b].p, base, m,
                              first_index + (uint32_t)rb.r0, t->k, t->view(), c->sm_count, c->stream[b]));
        c->launches++;
    }
    if ((rc = sync_all(c))) return rc;
    return table_refresh(t);
}

int tg_label_bundles_dev(tg_table* t, const void* d_recs, uint64_t nbytes, const void* d_offs, uint64_t nbundles,
                         uint32_t first_index) {
    if (!t || !d_recs || !d_offs) return fail(TG_ERR_ARG, "tg_label_bundles_dev: null argument");
    if (t->kind != TG_TABLE_LABEL) return fail(TG_ERR_ARG, "tg_label_bundles_dev needs a TG_TABLE_LABEL table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    CU(launch_label_tiles((const uint8_t*)d_recs, nbytes, (const uint64_t*)d_offs, 0, nbundles, first_index, t->k,
                          t->view(), c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}

static int ensure_lut(tg_ctx* c, const uint8_t* entropy_ok) {
    CU(c->lut.ensure(26 * 26 * 26));
    CU(cudaMemcpyAsync(c->lut.p, entropy_ok, 26 * 26 * 26, cudaMemcpyHostToDevice, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    return TG_OK;
}

int tg_assign_reads(tg_table* t, const char* recs, const uint64_t* offs, uint64_t nreads, int strand,
                    const uint8_t* entropy_ok, int32_t* best, int32_t* pct, int32_t* score) {
    if (!t || !entropy_ok || ((!recs || !offs || !best || !pct) && nreads))
        return fail(TG_ERR_ARG, "tg_assign_reads: null argument");
    if (t->kind != TG_TABLE_LABEL) return fail(TG_ERR_ARG, "tg_assign_reads needs a TG_TABLE_LABEL table");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (nreads == 0) return TG_OK;
    int rc;
    if ((rc = sync_all(c))) return rc;
    if ((rc = ensure_lut(c, entropy_ok))) return rc;
    if (t->k >= MIN_FAST_K && nreads <= 0x7FFFFFF0ull && is_held(c, recs, offs[nreads])) {
        bool done = false;
        if ((rc = held_pass(c, offs, nreads, &done, [&](const uint8_t* d, const uint64_t* d_offs, LongList ll) -> int {
                const uint32_t* ord = nullptr;
                if (int r3 = locus_order_async(c, 0, d, d_offs, 0, nreads, t->k, &ord, offs)) return r3;
                CU(launch_assign(d, d_offs, 0, nreads, t->k, strand, t->slots, t->g, (const uint8_t*)c->lut.p,
                                 (int32_t*)c->held.out_a.p, (int32_t*)c->held.out_b.p, (int32_t*)c->held.out_c.p, ll, ord, c->stream[0]));
                CU(launch_assign_long_auto(d, d_offs, 0, t->k, strand, t->slots, t->g, (const uint8_t*)c->lut.p,
                                           (int32_t*)c->held.out_a.p, (int32_t*)c->held.out_b.p, (int32_t*)c->held.out_c.p, ll,
                                           c->long_scratch.p, c->long_scratch_bytes, c->d_error, c->sm_count * 2, c->stream[0]));
                c->launches += 2;
                return TG_OK;
            }, best, pct, score))) return rc;
        if (done) return TG_OK;
    }
    const std::vector<ReadBatch> batches = split_reads(offs, nreads, c->batch_bytes);
    struct Pending { bool live = false; ReadBatch rb; } pend[2];
    auto drain = [&](int b) -> int {
        if (!pend[b].live) return TG_OK;
        const ReadBatch rb = pend[b].rb;
        const uint64_t m = rb.r1 - rb.r0, base = offs[rb.r0];
        int r2 = finish_long(c, b, t->k, assign_long_scratch_bytes,
            [&](unsigned n_long, unsigned max_win, void* scratch, int nctas) {
                return launch_assign_long((const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, t->k, strand,
                                          t->slots, t->g, (const uint8_t*)c->lut.p, (int32_t*)c->out_a[b].p,
                                          (int32_t*)c->out_b[b].p, (int32_t*)c->out_c[b].p,
                                          (const unsigned int*)c->long_idx[b].p, n_long, max_win, scratch, nctas,
                                          c->stream[b]);
            });
        if (r2) return r2;
        CU(cudaMemcpyAsync(best + rb.r0, c->out_a[b].p, m * 4, cudaMemcpyDeviceToHost, c->stream[b]));
        CU(cudaMemcpyAsync(pct + rb.r0, c->out_b[b].p, m * 4, cudaMemcpyDeviceToHost, c->stream[b]));
        if (score) CU(cudaMemcpyAsync(score + rb.r0, c->out_c[b].p, m * 4, cudaMemcpyDeviceToHost, c->stream[b]));
        CU(cudaStreamSynchronize(c->stream[b]));
        pend[b].live = false;
        return TG_OK;
    };
    for (size_t i = 0; i < batches.size(); i++) {
        const int b = (int)(i & 1);
        if ((rc = drain(b))) return rc;
        const ReadBatch rb = batches[i];
        const uint64_t m = rb.r1 - rb.r0, base = offs[rb.r0], nb = offs[rb.r1] - base;
        if ((rc = upload_records(c, b, recs + base, nb))) return rc;
        CU(c->offs[b].ensure((m + 1) * 8));
        CU(c->out_a[b].ensure(m * 4)); CU(c->out_b[b].ensure(m * 4)); CU(c->out_c[b].ensure(m * 4));
        CU(c->long_idx[b].ensure(m * 4));
        CU(cudaMemcpyAsync(c->offs[b].p, offs + rb.r0, (m + 1) * 8, cudaMemcpyHostToDevice, c->stream[b]));
        CU(cudaMemsetAsync(c->d_long_hdr[b], 0, 2 * sizeof(unsigned int), c->stream[b]));
        LongList ll{c->d_long_hdr[b], c->d_long_hdr[b] + 1, (unsigned int*)c->long_idx[b].p};
        if (t->k >= MIN_FAST_K) {
            const uint32_t* ord = nullptr;
            if ((rc = locus_order_async(c, b, (const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, m, t->k, &ord))) return rc;
            CU(launch_assign((const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, m, t->k, strand, t->slots,
                             t->g, (const uint8_t*)c->lut.p, (int32_t*)c->out_a[b].p, (int32_t*)c->out_b[b].p,
                             (int32_t*)c->out_c[b].p, ll, ord, c->stream[b]));
        } else {
            int nctas = 0; size_t need = 0;
            const unsigned max_win = batch_max_windows(offs, rb.r0, rb.r1, t->k);
            long_launch_shape(c, (unsigned)m, max_win, t->k, assign_long_scratch_bytes, &nctas, &need);
            CU(cudaStreamSynchronize(c->stream[b ^ 1]));
            CU(c->scratch.ensure(need));
            CU(launch_assign_long((const uint8_t*)c->recs[b].p, (const uint64_t*)c->offs[b].p, base, t->k, strand, t->slots,
                                  t->g, (const uint8_t*)c->lut.p, (int32_t*)c->out_a[b].p, (int32_t*)c->out_b[b].p,
                                  (int32_t*)c->out_c[b].p, nullptr, (unsigned)m, max_win, c->scratch.p, nctas, c->stream[b]));
        }
        c->launches++;
        pend[b].live = true; pend[b].rb = rb;
    }
    if ((rc = drain(0))) return rc;
    if ((rc = drain(1))) return rc;
    return TG_OK;
}

int tg_assign_reads_dev(tg_table* t, const void* d_recs, const void* d_offs, uint64_t nreads, int strand,
                        const void* d_entropy_ok, void* d_best, void* d_pct) {
    if (!t || !d_recs || !d_offs || !d_entropy_ok || !d_best || !d_pct)
        return fail(TG_ERR_ARG, "tg_assign_reads_dev: null argument");
    tg_ctx* c = t->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    if (nreads > 0x7FFFFFF0ull) return fail(TG_ERR_ARG, "tg_assign_reads_dev: at most 2^31 reads per call");
    if (t->k < MIN_FAST_K) return fail(TG_ERR_ARG, "tg_assign_reads_dev: k >= %d (shorter k-mers: the host-buffer entry point)", MIN_FAST_K);
    const int b = 0;
    CU(c->long_idx[b].ensure(nreads * 4));
    CU(cudaMemsetAsync(c->d_long_hdr[b], 0, 2 * sizeof(unsigned int), c->stream[b]));
    LongList ll{c->d_long_hdr[b], c->d_long_hdr[b] + 1, (unsigned int*)c->long_idx[b].p};
    CU(c->long_scratch.ensure(c->long_scratch_bytes));
    const uint32_t* ord = nullptr;
    if (int rc = locus_order_async(c, b, (const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, t->k, &ord)) return rc;
    CU(launch_assign((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, nreads, t->k, strand, t->slots, t->g,
                     (const uint8_t*)d_entropy_ok, (int32_t*)d_best, (int32_t*)d_pct, nullptr, ll, ord, c->stream[b]));
    CU(launch_assign_long_auto((const uint8_t*)d_recs, (const uint64_t*)d_offs, 0, t->k, strand, t->slots, t->g,
                               (const uint8_t*)d_entropy_ok, (int32_t*)d_best, (int32_t*)d_pct, nullptr, ll,
                               c->long_scratch.p, c->long_scratch_bytes, c->d_error, c->sm_count * 2, c->stream[b]));
    c->launches += 2;
    return TG_OK;
}

// compute_entropy(string&) of Chrysalis/analysis/sequenceUtil.cc:326-355, evaluated for every count tuple:
// fp32 throughout, slots in G,A,T,C order, log() on a float argument resolves to the float overload.
void tg_entropy_table(int k, float min_entropy, uint8_t* ok) {
    memset(ok, 0, 26 * 26 * 26);
    if (k > 25) k = 25;   // the table is dimensioned for k <= 25 (ReadsToTranscripts hard-codes k = 25)
    for (int g = 0; g <= k; g++)
        for (int a = 0; a + g <= k; a++)
            for (int tt = 0; tt + a + g <= k; tt++) {
                const int cnt[4] = {g, a, tt, k - g - a - tt};
                float entropy = 0;
                for (int i = 0; i < 4; i++) {
                    const float prob = (float)cnt[i] / (float)k;
                    if (prob > 0) {
                        const float val = prob * logf(1 / prob) / logf(2.0f);
                        entropy += val;
                    }
                }
                ok[(g * 26 + a) * 26 + tt] = !(entropy < min_entropy);
            }
}

// ---------------------------------------------------------------------------------------------------------
// GraphFromFasta weldmer counting
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace tg { unsigned long long weld_hash_host(unsigned long long lo, unsigned hi); }

struct tg_weld {
    tg_ctx* ctx = nullptr;
    int kk = 48;
    uint64_t n = 0, cap = 0;
    WeldSlot* d_slots = nullptr;
    std::vector<int64_t> slot_of;        // table slot of input weldmer i, -1 = unmatchable (a non-ACGT character)
};

extern "C" {

int tg_weld_create(tg_ctx* c, int kk, const char* weldmers, uint64_t n, tg_weld** out) {
    if (!c || !out || (!weldmers && n)) return fail(TG_ERR_ARG, "tg_weld_create: null argument");
    if (kk < 33 || kk > 48) return fail(TG_ERR_ARG, "weldmer length %d unsupported (33..48; GraphFromFasta's default -kk is 48)", kk);
    if (bind(c)) return TG_ERR_CUDA;
    tg_weld* w = new tg_weld();
    w->ctx = c; w->kk = kk; w->n = n;
    uint64_t cap = 1024;
    while (cap < 4 * n) cap <<= 1;                       // load <= 0.25: a miss settles on the first slot three times out of four
    w->cap = cap;
    std::vector<WeldSlot> host(cap);
    memset(host.data(), 0, cap * sizeof(WeldSlot));
    w->slot_of.assign(n, -1);
    for (uint64_t i = 0; i < n; i++) {
        const char* s = weldmers + i * (uint64_t)kk;
        unsigned long long p0 = 0, p1 = 0;
        bool ok = true;
        for (int b = 0; b < kk && ok; b++) {
            const unsigned ch = (unsigned char)s[b];
            if (!base_valid(ch)) { ok = false; break; }
            const unsigned code = base_code(ch);
            p0 |= (unsigned long long)(code & 1u) << b;
            p1 |= (unsigned long long)(code >> 1) << b;
        }
        if (!ok) continue;                               // can never equal a window of read bases
        const unsigned long long lo = p0 | (p1 << 48);
        const unsigned hi = (unsigned)(p1 >> 16);
        uint64_t at = weld_hash_host(lo, hi) & (cap - 1);
        while ((host[at].cnt & WELD_OCCUPIED) && !(host[at].lo == lo && host[at].hi == hi)) at = (at + 1) & (cap - 1);
        host[at].lo = lo; host[at].hi = hi; host[at].cnt = WELD_OCCUPIED;
        w->slot_of[i] = (int64_t)at;
    }
    cudaError_t e = cudaMalloc(&w->d_slots, cap * sizeof(WeldSlot));
    if (e != cudaSuccess) { delete w; CU(e); }
    CU(cudaMemcpyAsync(w->d_slots, host.data(), cap * sizeof(WeldSlot), cudaMemcpyHostToDevice, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    *out = w;
    return TG_OK;
}

void tg_weld_destroy(tg_weld* w) {
    if (!w) return;
    cudaSetDevice(w->ctx->device);
    cudaDeviceSynchronize();
    if (w->d_slots) cudaFree(w->d_slots);
    delete w;
}

int tg_weld_count_reads_dev(tg_weld* w, const void* d_recs, uint64_t nbytes) {
    if (!w || !d_recs) return fail(TG_ERR_ARG, "tg_weld_count_reads_dev: null argument");
    tg_ctx* c = w->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    CU(launch_weld_tiles((const uint8_t*)d_recs, nbytes, w->kk, w->d_slots, w->cap - 1, c->sm_count, c->stream[0]));
    c->launches++;
    return TG_OK;
}

int tg_weld_count_reads(tg_weld* w, const char* recs, uint64_t nbytes) {
    if (!w || (!recs && nbytes)) return fail(TG_ERR_ARG, "tg_weld_count_reads: null argument");
    tg_ctx* c = w->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc;
    uint64_t pos = 0;
    for (int it = 0; pos < nbytes; it++) {
        const int b = it & 1;
        const uint64_t end = batch_end(recs, pos, nbytes, c->batch_bytes);
        const uint64_t n = end - pos;
        CU(cudaStreamSynchronize(c->stream[b]));          // staging buffer b is free again
        if ((rc = upload_records(c, b, recs + pos, n))) return rc;
        CU(launch_weld_tiles((const uint8_t*)c->recs[b].p, n, w->kk, w->d_slots, w->cap - 1, c->sm_count, c->stream[b]));
        c->launches++;
        pos = end;
    }
    return sync_all(c);
}

int tg_weld_counts(tg_weld* w, int32_t* counts) {
    if (!w || (!counts && w->n)) return fail(TG_ERR_ARG, "tg_weld_counts: null argument");
    tg_ctx* c = w->ctx;
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    if (rc) return rc;
    std::vector<WeldSlot> host(w->cap);
    CU(cudaMemcpy(host.data(), w->d_slots, w->cap * sizeof(WeldSlot), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < w->n; i++)
        counts[i] = w->slot_of[i] < 0 ? 0 : (int32_t)(host[(size_t)w->slot_of[i]].cnt & ~WELD_OCCUPIED);
    return TG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// device memory + timing + measurement utilities
// ---------------------------------------------------------------------------------------------------------
int tg_dev_alloc(tg_ctx* c, uint64_t bytes, void** dptr) {
    if (!c || !dptr) return fail(TG_ERR_ARG, "tg_dev_alloc: null argument");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaMalloc(dptr, bytes ? bytes : 1));
    return TG_OK;
}

int tg_dev_records_alloc(tg_ctx* c, uint64_t nbytes, void** dptr) {
    if (!c || !dptr) return fail(TG_ERR_ARG, "tg_dev_records_alloc: null argument");
    if (bind(c)) return TG_ERR_CUDA;
    const uint64_t padded = padded_record_bytes(nbytes);
    CU(cudaMalloc(dptr, padded));
    CU(cudaMemsetAsync((char*)*dptr + nbytes, '\n', padded - nbytes, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    return TG_OK;
}

int tg_dev_free(tg_ctx* c, void* dptr) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    if (dptr) CU(cudaFree(dptr));
    return TG_OK;
}

int tg_memcpy_h2d(tg_ctx* c, void* dptr, const void* host, uint64_t bytes) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    return TG_OK;
}

int tg_memcpy_d2h(tg_ctx* c, void* host, const void* dptr, uint64_t bytes) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    return TG_OK;
}

int tg_kernel_times(tg_ctx* c, char* out, uint64_t out_bytes) {
    if (!c || !out || out_bytes == 0) return fail(TG_ERR_ARG, "tg_kernel_times: bad argument");
    if (bind(c)) return TG_ERR_CUDA;
    int rc = sync_all(c);
    if (rc) return rc;
    struct Acc { const char* name; double ms; unsigned n; };
    std::vector<Acc> acc;
    for (auto& sp : c->timer.spans) {
        float ms = 0;
        if (sp.a && sp.b && cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            bool found = false;
            for (auto& a : acc) if (!strcmp(a.name, sp.name)) { a.ms += ms; a.n++; found = true; break; }
            if (!found) acc.push_back({sp.name, ms, 1});
        }
        if (sp.a) cudaEventDestroy(sp.a);
        if (sp.b) cudaEventDestroy(sp.b);
    }
    cudaGetLastError();
    c->timer.spans.clear();
    std::string txt;
    char line[256];
    for (auto& a : acc) { snprintf(line, sizeof line, "%s\t%.6f\t%u\n", a.name, a.ms, a.n); txt += line; }
    if (txt.size() + 1 > out_bytes) return fail(TG_ERR_ARG, "tg_kernel_times: buffer too small (%zu bytes needed)", txt.size() + 1);
    memcpy(out, txt.c_str(), txt.size() + 1);
    return TG_OK;
}

int tg_memcpy_d2d(tg_ctx* c, void* dst, const void* src, uint64_t bytes) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream[0]));
    CU(cudaStreamSynchronize(c->stream[0]));
    return TG_OK;
}

int tg_memset_dev(tg_ctx* c, void* dst, int value, uint64_t bytes) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaMemsetAsync(dst, value, bytes, c->stream[0]));
    return TG_OK;
}

int tg_timer_start(tg_ctx* c) {
    if (!c) return fail(TG_ERR_ARG, "null ctx");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaEventRecord(c->t0, c->stream[0]));
    return TG_OK;
}

int tg_timer_stop(tg_ctx* c, float* ms) {
    if (!c || !ms) return fail(TG_ERR_ARG, "null argument");
    if (bind(c)) return TG_ERR_CUDA;
    CU(cudaEventRecord(c->t1, c->stream[0]));
    CU(cudaEventSynchronize(c->t1));
    CU(cudaEventElapsedTime(ms, c->t0, c->t1));
    return TG_OK;
}

int tg_gups(tg_ctx* c, uint64_t slots, uint64_t nops, int mode, int reps, float* best_ms) {
    if (!c || !best_ms || slots == 0) return fail(TG_ERR_ARG, "tg_gups: bad argument");
    if (mode < 0 || mode > 2) return fail(TG_ERR_ARG, "tg_gups: mode must be 0, 1 or 2");
    if (bind(c)) return TG_ERR_CUDA;
    Slot* p = nullptr;
    unsigned long long* sink = nullptr;
    CU(cudaMalloc(&p, slots * sizeof(Slot)));
    CU(cudaMalloc(&sink, 8));
    CU(cudaMemsetAsync(p, 0, slots * sizeof(Slot), c->stream[0]));
    float best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {   // first repetition is the warm-up
        CU(cudaEventRecord(c->t0, c->stream[0]));
        CU(launch_gups(p, slots, nops, mode, sink, c->sm_count, c->stream[0]));
        c->launches++;
        CU(cudaEventRecord(c->t1, c->stream[0]));
        CU(cudaEventSynchronize(c->t1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->t0, c->t1));
        if (r > 0 && ms < best) best = ms;
    }
    cudaFree(p); cudaFree(sink);
    *best_ms = best;
    return TG_OK;
}

int tg_synth_reads_dev(tg_ctx* c, const char* tx, const uint64_t* tx_offs, const uint64_t* tx_cum, uint32_t ntx,
                       uint64_t npairs, int read_len, int frag_mean, int frag_sd, uint32_t err_per_million,
                       uint32_t n_per_million, uint64_t seed, int stranded, void* d_recs) {
    if (!c || !tx || !tx_offs || !tx_cum || !d_recs || ntx == 0) return fail(TG_ERR_ARG, "tg_synth_reads_dev: bad argument");
    if (bind(c)) return TG_ERR_CUDA;
    uint8_t* d_tx = nullptr; uint64_t* d_offs = nullptr; uint64_t* d_cum = nullptr;
    const uint64_t txb = tx_offs[ntx];
    CU(cudaMalloc(&d_tx, txb ? txb : 1));
    CU(cudaMalloc(&d_offs, (ntx + 1) * 8));
    CU(cudaMalloc(&d_cum, ntx * 8));
    CU(cudaMemcpyAsync(d_tx, tx, txb, cudaMemcpyHostToDevice, c->stream[0]));
    CU(cudaMemcpyAsync(d_offs, tx_offs, (ntx + 1) * 8, cudaMemcpyHostToDevice, c->stream[0]));
    CU(cudaMemcpyAsync(d_cum, tx_cum, ntx * 8, cudaMemcpyHostToDevice, c->stream[0]));
    CU(launch_synth_reads(d_tx, d_offs, d_cum, ntx, npairs, read_len, frag_mean, frag_sd, err_per_million,
                          n_per_million, seed, stranded, (uint8_t*)d_recs, c->stream[0]));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream[0]));
    cudaFree(d_tx); cudaFree(d_offs); cudaFree(d_cum);
    return TG_OK;
}

}  // extern "C"
