// Several GPUs behind one drop-in executable ("reads shard by input offset").  The per-read tools
// (fastaToKmerCoverageStats, ReadsToTranscripts) are embarrassingly parallel over reads once every GPU holds the table:
// TRINITY_GPUS=0,1,...,7 makes the tool open one context per listed device, REPLICATE the (read-only) table on each, and cut
// every batch of reads into contiguous ranges -- one per GPU, balanced by bytes -- that are processed by one host thread
// each and written back into the batch's result arrays at the reads' own positions, so the output is byte-identical to a
// single-GPU run.  Nothing but the environment selects it: argv stays the reference's.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "tg_loader.hpp"

namespace tgh {

// device list: TRINITY_GPUS="0,1,2" (several), else TRINITY_GPU=<index> (one), else device 0
inline std::vector<int> device_list() {
    std::vector<int> devs;
    if (const char* e = getenv("TRINITY_GPUS")) {
        const char* p = e;
        while (*p) {
            char* end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            devs.push_back((int)v);                 // (a device may be listed twice: two contexts on it -- the tests do that on a one-GPU box)
            p = end;
            while (*p == ',' || *p == ' ') p++;
        }
    }
    if (devs.empty()) devs.push_back(getenv("TRINITY_GPU") ? atoi(getenv("TRINITY_GPU")) : 0);
    return devs;
}

struct GpuSet {
    std::vector<tg_ctx*> ctx;
    void open() {
        for (int d : device_list()) {
            tg_ctx* c = nullptr;
            if (tg_init(d, &c) != TG_OK) die(3, "tg_init");
            ctx.push_back(c);
        }
    }
    size_t size() const { return ctx.size(); }
    void close() { for (tg_ctx* c : ctx) tg_destroy(c); ctx.clear(); }
};

// f(g) on one host thread per GPU (g = 0 runs on the caller's thread); the first failure is reported like TGC does.
// (tg_last_error is thread-local: a worker keeps its own message.)
template <typename F>
inline void on_every_gpu(size_t n, F f) {
    std::vector<std::string> errs(n);
    std::vector<int> rcs(n, 0);
    auto run = [&](size_t g) { rcs[g] = f(g); if (rcs[g] != TG_OK) errs[g] = tg_last_error(); };
    std::vector<std::thread> th;
    for (size_t g = 1; g < n; g++) th.emplace_back(run, g);
    run(0);
    for (auto& t : th) t.join();
    for (size_t g = 0; g < n; g++)
        if (rcs[g] != TG_OK) { fprintf(stderr, "ERROR: GPU %zu of %zu: %s\n", g, n, errs[g].c_str()); exit(3); }
}

// reads [0, n) of a batch (offs[n + 1] = record starts) -> `parts` contiguous ranges with about the same number of bytes
inline std::vector<std::pair<size_t, size_t>> split_reads_by_bytes(const uint64_t* offs, size_t n, size_t parts) {
    std::vector<std::pair<size_t, size_t>> out;
    if (parts < 1) parts = 1;
    const uint64_t total = n ? offs[n] - offs[0] : 0;
    size_t a = 0;
    for (size_t p = 0; p < parts; p++) {
        size_t b = n;
        if (p + 1 < parts) {
            const uint64_t want = offs[0] + total * (p + 1) / parts;
            b = (size_t)(std::lower_bound(offs + a, offs + n + 1, want) - offs);
            if (b > n) b = n;
            if (b < a) b = a;
        }
        out.emplace_back(a, b);
        a = b;
    }
    return out;
}

}  // namespace tgh
