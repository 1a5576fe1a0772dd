// Parallel, order-preserving FASTA parsing for the drop-in executables.
//
// Once the k-mer work runs on the GPU the tools are bound by their host side: one thread walking a multi-gigabyte FASTA
// (SURVEY §7 "host side becomes the bottleneck").  A FASTA splits cleanly wherever a line starts with '>': the file is cut
// into chunks at such lines, a pool of threads parses the chunks with the SAME reader the serial path uses (so every
// quirk of the reference reader is reproduced by construction -- a chunk is itself a FASTA text), and the consumer takes
// the parsed batches strictly in file order.  Only a window of chunks is in flight, so memory stays bounded whatever the
// file size.
#pragma once
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "fasta_io.hpp"

namespace tgio {

// chunk boundaries: 0 = b[0] < b[1] < ... < b[n] = size, every inner boundary at a '>' that starts a line
// after_sequence_line: only cut at a '>' line whose PREVIOUS line does not start with '>'.  Chrysalis' DNAStringStreamFast
// takes the line after a header as sequence whatever it starts with (DNAVector.cc:1456-1501), so of two consecutive '>'
// lines the second may be sequence; a '>' line behind a plain line is a header for every reader.
inline std::vector<size_t> fasta_chunks(const char* data, size_t size, size_t target, bool after_sequence_line = false) {
    std::vector<size_t> b{0};
    if (target == 0) target = 1;
    size_t pos = target;
    const char* end = data + size;
    auto cut_ok = [&](const char* p) {          // p: a line start holding '>'
        if (!after_sequence_line) return true;
        if (p - data < 2) return false;         // (the first line has no previous line; position 0 is a boundary anyway)
        const char* prev = (const char*)memrchr(data, '\n', (size_t)(p - 1 - data));      // newline before the previous line
        const char* ps = prev ? prev + 1 : data;
        return *ps != '>';
    };
    while (pos < size) {
        // next line start at or after pos that begins with '>'
        const char* p = data + pos;
        if (!(pos > 0 && data[pos - 1] == '\n' && *p == '>' && cut_ok(p))) {
            for (;;) {
                const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
                if (!nl || nl + 1 >= end) { p = end; break; }
                p = nl + 1;
                if (*p == '>' && cut_ok(p)) break;
            }
        }
        const size_t at = (size_t)(p - data);
        if (at >= size) break;
        b.push_back(at);
        pos = at + target;
    }
    b.push_back(size);
    return b;
}

// Parses chunk i with `parse(data + b[i], b[i+1] - b[i], batch)` on a pool of threads; next() hands the batches out in
// order.  At most `window` parsed-but-unconsumed batches exist at a time.
class OrderedChunkParser {
public:
    using ParseFn = std::function<void(const char*, size_t, RecordBatch&)>;
    OrderedChunkParser(const char* data, size_t size, size_t target_chunk, unsigned threads, unsigned window, ParseFn parse,
                       bool after_sequence_line = false)
        : data_(data), bounds_(fasta_chunks(data, size, target_chunk, after_sequence_line)), window_(window < 2 ? 2 : window), parse_(std::move(parse)),
          slots_(window_), ready_(window_, 0) {
        const size_t n = nchunks();
        if (threads < 1) threads = 1;
        if ((size_t)threads > n) threads = (unsigned)(n ? n : 1);
        for (unsigned t = 0; t < threads; t++) workers_.emplace_back([this] { work(); });
    }
    ~OrderedChunkParser() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    size_t nchunks() const { return bounds_.size() - 1; }
    // the next batch in file order (moved out), false at the end
    bool next(RecordBatch& out) {
        if (consumed_ >= nchunks()) return false;
        std::unique_lock<std::mutex> g(m_);
        const size_t s = consumed_ % window_;
        cv_.wait(g, [&] { return ready_[s] != 0; });
        out = std::move(slots_[s]);
        slots_[s] = RecordBatch();
        ready_[s] = 0;
        consumed_++;
        g.unlock();
        cv_.notify_all();
        return true;
    }
private:
    void work() {
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> g(m_);
                i = issued_;
                if (stop_ || i >= nchunks()) return;
                issued_++;
                cv_.wait(g, [&] { return stop_ || i < consumed_ + window_; });     // the slot of chunk i is free
                if (stop_) return;
            }
            RecordBatch rb;
            parse_(data_ + bounds_[i], bounds_[i + 1] - bounds_[i], rb);
            {
                std::lock_guard<std::mutex> g(m_);
                slots_[i % window_] = std::move(rb);
                ready_[i % window_] = 1;
            }
            cv_.notify_all();
        }
    }
    const char* data_;
    std::vector<size_t> bounds_;
    size_t window_;
    ParseFn parse_;
    std::vector<RecordBatch> slots_;
    std::vector<char> ready_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    size_t issued_ = 0, consumed_ = 0;
    bool stop_ = false;
};

// run fn(t) for t = 0..n-1 on n threads (the caller's thread takes t = 0)
template <typename Fn>
inline void parallel_for_threads(unsigned n, Fn fn) {
    std::vector<std::thread> ts;
    for (unsigned t = 1; t < n; t++) ts.emplace_back([&fn, t] { fn(t); });
    fn(0);
    for (auto& th : ts) th.join();
}

inline unsigned host_threads(unsigned cap) {
    unsigned h = std::thread::hardware_concurrency();
    if (h == 0) h = 1;
    if (const char* e = getenv("TRINITY_GPU_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) h = (unsigned)v; }
    return h < cap ? h : cap;
}

}  // namespace tgio
