// fastaToKmerCoverageStats -- drop-in for Inchworm/bin/fastaToKmerCoverageStats (reference:
// Inchworm/src/fastaToKmerCoverageStats.cpp).  Same argv, same stdout format, same exit codes; the k-mer table
// and the per-read statistics run on the GPU through libtrinity_gpu.
#include <math.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <map>
#include <memory>
#include <thread>
#include <string>
#include <vector>

#include "fasta_io.hpp"
#include "fmt_float.hpp"
#include "par_fasta.hpp"
#include "multi_gpu.hpp"
#include "tg_loader.hpp"
#include "tg_sidecar.hpp"

using namespace tgio;

static const int MAX_THREADS = 6;   // kept only for the usage text / CLI compatibility
static const unsigned HOST_THREADS_CAP = 32;      // parse / format threads (TRINITY_GPU_HOST_THREADS overrides the core count)
static const size_t CHUNK_BYTES = 48u << 20;        // FASTA bytes per parsed batch

// Inchworm ArgProcessor (Inchworm/src/argProcessor.cpp:5-23): every token starting with '-' is a flag and the
// following token, whatever it is, is recorded as its value.
struct Args {
    std::map<std::string, std::string> val;
    std::map<std::string, bool> set;
    Args(int argc, char** argv) {
        for (int i = 1; i < argc; i++)
            if (argv[i][0] == '-') {
                set[argv[i]] = true;
                if (i != argc - 1) val[argv[i]] = argv[i + 1];
            }
    }
    bool isSet(const char* a) const { return set.count(a) != 0; }
    std::string str(const char* a) { return val[a]; }
    int i(const char* a) { return atoi(val[a].c_str()); }
};

static void usage() {
    fprintf(stderr,
            "\n\nUsage: \n"
            "  --reads  <str>             :fasta file containing target reads for kmer coverage stats\n"
            "\n and source of kmers: \n"
            "  --kmers  <str>             :fasta file containing kmers\n"
            "      or \n"
            "  --kmers_from_reads <str>   :fasta file containing reads as source of kmers\n"
            "\n* optional:\n"
            "  --kmer_size <int>          :default = 25\n"
            "  --DS                             :double-stranded RNA-Seq mode (not strand-specific)\n"
            "  --capture_coverage_info    :writes coverage info file.\n"
            "  --monitor <int>            :verbose level for debugging\n"
            "  --num_threads <int>        :number of threads\n"
            "\n\n\n");
}

// TRINITY_GPU_TRACE=1: wall-clock seconds of the tool's phases on stderr (where does a whole-process run spend its time?)
struct Trace {
    bool on = getenv("TRINITY_GPU_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double acc[4] = {0, 0, 0, 0};                 // parse wait, GPU call, format/write wait, other
    double now() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    void mark(const char* what) const { if (on) fprintf(stderr, "[trace] %8.3f s  %s\n", now(), what); }
    // absolute time, to place the process's start and end inside the caller's own clock (exec + exit are not ours to see)
    void wall(const char* what) const {
        if (!on) return;
        struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts);
        fprintf(stderr, "[trace] wall %.3f  %s\n", (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec, what);
    }
};

int main(int argc, char** argv) {
    Args args(argc, argv);
    if (args.isSet("--help") || !(args.isSet("--reads") && (args.isSet("--kmers") || args.isSet("--kmers_from_reads")))) {
        usage();
        return 1;
    }
    const std::string reads_file = args.str("--reads");
    const bool is_DS = !args.isSet("--SS");           // "--DS" is accepted and ignored, like the reference (:75)
    int K = 25;
    if (args.isSet("--kmer_size")) {
        K = args.i("--kmer_size");
        if (K < 20) { fprintf(stderr, "Error, min kmer size is 20"); return 2; }
    }
    if (K > 32) { fprintf(stderr, "ERROR encountered: \n\nKmer length exceeds max of 32"); return 1; }
    const bool capture = args.isSet("--capture_coverage_info");

    Trace trace;
    trace.wall("main");
    tgh::GpuSet gpus;                       // TRINITY_GPUS=0,1,..: the reads of every batch are split over these devices
    // Creating the CUDA context(s) takes a few tenths of a second in which the host has nothing to wait for: it runs on a
    // thread of its own while the first chunks of the input are being parsed, and is joined before the first library call.
    struct Opener {
        std::thread th; bool joined = false;
        void wait() { if (!joined) { th.join(); joined = true; } }
        ~Opener() { wait(); }
    } opener;
    opener.th = std::thread([&gpus] { gpus.open(); });
    tg_ctx* ctx = nullptr;                  // the table is built on the first device and replicated (multi_gpu.hpp)
    auto wait_ctx = [&] {
        if (opener.joined) return;
        opener.wait();
        trace.mark("device context(s) open");
        ctx = gpus.ctx[0];
    };
    tg_table* table = nullptr;
    std::string err;
    const unsigned nthreads_host = host_threads(HOST_THREADS_CAP);
    const unsigned parse_window = nthreads_host + 2;      // parsed-or-in-flight chunks: every thread of the pool has one

    FileView rv;
    std::unique_ptr<OrderedChunkParser> reads_parser;
    auto start_reads_parser = [&]() -> bool {
        if (reads_parser) return true;
        if (!rv.open(reads_file, &err)) { fprintf(stderr, "ERROR encountered: \n\nError, %s", err.c_str()); return false; }
        reads_parser.reset(new OrderedChunkParser(rv.data, rv.size, CHUNK_BYTES, nthreads_host, parse_window,
            [](const char* d, size_t n, RecordBatch& rb) {
                rb.recs.reserve(n); rb.offs.reserve(n / 48 + 16); rb.names.reserve(n / 6); rb.name_offs.reserve(n / 48 + 16);
                InchwormFastaReader rd(d, n);
                const char* h; size_t hl;
                while (true) {
                    const size_t before = rb.recs.size();
                    if (!rd.next(&h, &hl, rb.recs)) break;                       // the cleaned sequence lands in the batch directly
                    if (rb.recs.size() == before) continue;                      // :132-133
                    const char* acc; size_t al;
                    accession_of(h, hl, &acc, &al);
                    rb.end_record();
                    rb.add_name(acc, al);
                }
            }));
        return true;
    };

    // ---- load the k-mer table --------------------------------------------------------------------------
    if (args.isSet("--kmers")) {
        // populate_kmer_counter_from_kmers (:181-228): `>COUNT\nKMER` records; stop at the first record with an
        // empty sequence; wrong-length k-mers are reported and skipped; count = atoi(header) as unsigned int
        FileView fv;
        if (!fv.open(args.str("--kmers"), &err)) { fprintf(stderr, "ERROR encountered: \n\nError, %s", err.c_str()); return 1; }
        fprintf(stderr, "-reading Kmer occurrences...\n");
        time_t start = time(NULL);
        wait_ctx();
        TGC(tg_table_create(ctx, TG_TABLE_COUNT, K, fv.size / 32 + 1024, &table));
        // Binary hand-off (tg_sidecar.hpp): if our `jellyfish dump` left `<kmers>.tgk` and it provably describes THIS
        // file (length + content hash), load the packed pairs and skip the text parse.  Same records, same
        // `table[canon] += count` as populate_kmer_counter_from_kmers, same diagnostics.
        bool from_sidecar = false;
        if (!tgside::disabled()) {
            FileView sv;
            std::string serr;
            if (sv.open(args.str("--kmers") + ".tgk", &serr) && sv.size >= sizeof(tgside::TgkHeader)) {
                const tgside::TgkHeader* sh = (const tgside::TgkHeader*)sv.data;
                if (memcmp(sh->magic, tgside::TGK_MAGIC, 8) == 0 && sh->k == (uint32_t)K && sh->text_bytes == fv.size &&
                    sv.size == sizeof(tgside::TgkHeader) + sh->n * 12) {
                    tgside::TextHash th;
                    th.update(fv.data, fv.size);
                    if (th.digest() == sh->text_hash) {
                        const uint64_t* sk = (const uint64_t*)(sv.data + sizeof(tgside::TgkHeader));
                        const uint32_t* sc = (const uint32_t*)(sv.data + sizeof(tgside::TgkHeader) + sh->n * 8);
                        const uint64_t STEP = 32u << 20;
                        for (uint64_t i = 0; i < sh->n; i += STEP)
                            TGC(tg_table_load_pairs(table, sk + i, sc + i, std::min<uint64_t>(STEP, sh->n - i), is_DS));
                        uint64_t cap = 0, distinct = 0;
                        TGC(tg_table_info(table, &cap, &distinct));
                        fprintf(stderr, "\n done parsing %lu Kmers, %llu added, taking %ld seconds.\n", (unsigned long)sh->n,
                                (unsigned long long)distinct, (long)(time(NULL) - start));
                        from_sidecar = true;
                    }
                }
            }
        }
        InchwormFastaReader rd(fv.data, from_sidecar ? 0 : fv.size);      // nothing left to parse after a sidecar load
        std::vector<uint64_t> keys; std::vector<uint32_t> vals;
        std::vector<char> seq;
        const size_t FLUSH = 8u << 20;
        keys.reserve(FLUSH); vals.reserve(FLUSH);
        unsigned long parsed = 0;
        const char* h; size_t hl;
        for (;;) {
            seq.clear();
            if (!rd.next(&h, &hl, seq)) break;
            if (seq.empty()) break;
            parsed++;
            if (seq.size() != (size_t)K) {
                fprintf(stderr, "ERROR: kmer %.*s is not of length: %d\n", (int)seq.size(), seq.data(), K);
                continue;
            }
            uint64_t key;
            if (!tgh::pack_kmer(seq.data(), K, &key)) {
                // kmer_to_intval throws on a non-GATC character (sequenceUtil.cpp:276-281) and the tool dies
                fprintf(stderr, "\n\nerror, kmer contains nongatc: %.*s", K, seq.data());
                return 1;
            }
            std::string hs(h, hl);
            keys.push_back(key);
            vals.push_back((uint32_t)atoi(hs.c_str()));
            if (keys.size() == FLUSH) {
                TGC(tg_table_load_pairs(table, keys.data(), vals.data(), keys.size(), is_DS));
                keys.clear(); vals.clear();
            }
        }
        if (!keys.empty()) TGC(tg_table_load_pairs(table, keys.data(), vals.data(), keys.size(), is_DS));
        if (!from_sidecar) {
            uint64_t cap = 0, distinct = 0;
            TGC(tg_table_info(table, &cap, &distinct));
            fprintf(stderr, "\n done parsing %lu Kmers, %llu added, taking %ld seconds.\n", parsed,
                    (unsigned long long)distinct, (long)(time(NULL) - start));
        }
    } else {
        // populate_kmer_counter_from_reads (:230-293): reads shorter than K+1 are skipped entirely
        FileView fv;
        if (!fv.open(args.str("--kmers_from_reads"), &err)) { fprintf(stderr, "ERROR encountered: \n\nError, %s", err.c_str()); return 1; }
        fprintf(stderr, "-storing Kmers...\n");
        // chunks of the file are parsed by a pool of threads and counted in file order (par_fasta.hpp)
        OrderedChunkParser parser(fv.data, fv.size, CHUNK_BYTES, nthreads_host, parse_window,
            [K](const char* d, size_t n, RecordBatch& rb) {
                rb.recs.reserve(n); rb.offs.reserve(n / 48 + 16);      // (no regrowth copies: the cleaned text is shorter than the chunk)
                InchwormFastaReader rd(d, n);
                const char* h; size_t hl;
                while (true) {
                    const size_t before = rb.recs.size();
                    if (!rd.next(&h, &hl, rb.recs)) break;          // the cleaned sequence lands in the batch directly
                    if (rb.recs.size() - before < (size_t)K + 1) { rb.recs.resize(before); continue; }
                    rb.end_record();
                }
            });
        wait_ctx();
        TGC(tg_table_create(ctx, TG_TABLE_COUNT, K, fv.size / 4 + 1024, &table));
        RecordBatch rb;
        while (parser.next(rb))
            if (!rb.recs.empty()) TGC(tg_count_reads(table, rb.recs.data(), rb.recs.size(), is_DS));
    }

    trace.mark("k-mer table loaded / counted");
    // ---- several GPUs: every device gets a replica of the (now read-only) table ----------------------------------
    std::vector<tg_table*> tables(1, table);
    if (gpus.size() > 1) {
        uint64_t cap = 0, distinct = 0, n = 0;
        uint64_t* keys = nullptr; uint32_t* vals = nullptr;
        TGC(tg_table_info(table, &cap, &distinct));
        TGC(tg_table_export(table, 0, 0xFFFFFFFFu, 0, 0, &keys, &vals, &n));
        tables.resize(gpus.size(), nullptr);
        tgh::on_every_gpu(gpus.size(), [&](size_t g) -> int {
            if (g == 0) return TG_OK;
            int rc = tg_table_create(gpus.ctx[g], TG_TABLE_COUNT, K, distinct + 1024, &tables[g]);
            const uint64_t STEP = 32u << 20;
            for (uint64_t i = 0; rc == TG_OK && i < n; i += STEP)
                rc = tg_table_load_pairs(tables[g], keys + i, vals + i, std::min<uint64_t>(STEP, n - i), is_DS);
            return rc;
        });
        tg_free(keys); tg_free(vals);
    }

    // ---- per-read statistics ----------------------------------------------------------------------------
    if (!start_reads_parser()) return 1;
    OrderedChunkParser& parser = *reads_parser;
    time_t start_time = time(NULL);
    OutBuf out(1);
    out.put("acc\tmedian_cov\tmean_cov\tstdev\ttid\n");
    // Pipeline (all in file order): a pool of threads parses chunks of the file; the main thread runs the statistics of
    // batch i on the GPU while the lines of batch i - 1 are being formatted by the pool's cores, and writes them out.
    struct Job {
        RecordBatch rb;
        std::vector<uint32_t> median, per_kmer;
        std::vector<float> mean, stdev;
        std::vector<std::vector<char>> bufs;
        std::vector<char> negs;
        std::thread formatter;
        uint64_t ticket = 0;                      // position of the batch in file order
    };
    constexpr unsigned NJOBS = 4;                 // batches between the GPU call and the last byte written
    Job jobs[NJOBS];
    bool negative = false;
    // Lines leave in file order, but not through the main thread: the thread that formatted batch i also writes it, as
    // soon as batch i - 1 is out (a ticket), while the main thread is already running batch i + 1 on the GPU.
    std::mutex out_mu;
    std::condition_variable out_cv;
    uint64_t out_next = 0;
    bool out_failed = false;
    auto format_job = [&](Job& jb) {
        const RecordBatch& rb = jb.rb;
        const size_t n = rb.count();
        const size_t nt = n < 4096 ? 1 : nthreads_host;
        jb.bufs.assign(nt, std::vector<char>());
        jb.negs.assign(nt, 0);
        // two %g conversions per read are the most expensive thing left on the host once the statistics come from the GPU
        parallel_for_threads((unsigned)nt, [&](unsigned w) {
            std::vector<char>& buf = jb.bufs[w];
            const size_t a = n * w / nt, b = n * (w + 1) / nt;
            buf.reserve((b - a) * 48);
            char num[64];
            auto put = [&](const char* p, size_t m) { buf.insert(buf.end(), p, p + m); };
            auto put_uint = [&](uint32_t v) { char t[12]; int m = 0; do { t[m++] = (char)('0' + v % 10); v /= 10; } while (v); while (m) buf.push_back(t[--m]); };
            for (size_t i = a; i < b; i++) {
                put(rb.name(i), rb.name_len(i));
                buf.push_back('\t'); put_uint(jb.median[i]);
                buf.push_back('\t'); put(num, (size_t)fmt_float(num, jb.mean[i]));
                buf.push_back('\t'); put(num, (size_t)fmt_float(num, jb.stdev[i]));
                put("\tthread:0", 9);
                if (capture) {
                    buf.push_back('\t');
                    const size_t L = rb.seq_len(i);
                    const size_t nw = L >= (size_t)K ? L - K + 1 : 0;
                    for (size_t j = 0; j < nw; j++) {
                        put_uint(jb.per_kmer[rb.offs[i] + j]);
                        if (j != nw - 1) buf.push_back(',');
                    }
                }
                buf.push_back('\n');
                if (jb.mean[i] < 0) jb.negs[w] = 1;
            }
        });
    };
    auto write_job = [&](Job& jb) {            // (on the job's own thread) its lines, once every earlier batch is out
        std::unique_lock<std::mutex> lk(out_mu);
        out_cv.wait(lk, [&] { return out_next == jb.ticket; });
        lk.unlock();
        for (size_t w = 0; w < jb.bufs.size(); w++) {
            const char* p = jb.bufs[w].data();
            size_t n = jb.bufs[w].size();
            while (n) {
                const ssize_t wr = ::write(1, p, n);
                if (wr < 0) { out_failed = true; break; }
                p += wr; n -= (size_t)wr;
            }
        }
        lk.lock();
        out_next = jb.ticket + 1;
        lk.unlock();
        out_cv.notify_all();
    };
    auto finish_job = [&](Job& jb) {           // wait until its lines are formatted and written
        if (!jb.formatter.joinable()) return;
        jb.formatter.join();
        for (size_t w = 0; w < jb.negs.size(); w++)
            if (jb.negs[w]) negative = true;
        jb.bufs.clear();
    };
    if (!out.flush()) { fprintf(stderr, "ERROR: write to stdout failed\n"); return 1; }     // the header line goes first
    uint64_t next_ticket = 0;
    for (unsigned it = 0;; it++) {
        Job& jb = jobs[it % NJOBS];
        double tt = trace.now();
        finish_job(jb);                          // the job that used this slot NJOBS batches ago
        trace.acc[2] += trace.now() - tt; tt = trace.now();
        if (!parser.next(jb.rb)) break;
        trace.acc[0] += trace.now() - tt;
        const size_t n = jb.rb.count();
        if (n == 0) continue;
        jb.median.resize(n); jb.mean.resize(n); jb.stdev.resize(n);
        if (capture) jb.per_kmer.assign(jb.rb.recs.size(), 0);
        tt = trace.now();
        if (gpus.size() == 1) {
            TGC(tg_cov_stats(table, jb.rb.recs.data(), jb.rb.offs.data(), n, is_DS, jb.median.data(), jb.mean.data(), jb.stdev.data(),
                             capture ? jb.per_kmer.data() : nullptr));
        } else {
            // contiguous ranges of the batch's reads, one per GPU; results land at the reads' own positions
            const auto ranges = tgh::split_reads_by_bytes(jb.rb.offs.data(), n, gpus.size());
            tgh::on_every_gpu(gpus.size(), [&](size_t g) -> int {
                const size_t a = ranges[g].first, b = ranges[g].second;
                if (a == b) return TG_OK;
                return tg_cov_stats(tables[g], jb.rb.recs.data(), jb.rb.offs.data() + a, b - a, is_DS, jb.median.data() + a,
                                    jb.mean.data() + a, jb.stdev.data() + a, capture ? jb.per_kmer.data() : nullptr);
            });
        }
        trace.acc[1] += trace.now() - tt;
        for (size_t i = 0; i < n; i++)
            if (jb.rb.seq_len(i) < (size_t)K)     // compute_kmer_coverage :305-310 (note the missing space, as in the reference)
                fprintf(stderr, "Sequence: %.*sis smaller than %d base pairs, skipping\n", (int)jb.rb.seq_len(i), jb.rb.seq(i), K);
        jb.ticket = next_ticket++;
        jb.formatter = std::thread([&format_job, &write_job, &jb] { format_job(jb); write_job(jb); });
    }
    for (unsigned j = 0; j < NJOBS; j++) finish_job(jobs[j]);
    if (trace.on) fprintf(stderr, "[trace] statistics loop: waiting for parsed batches %.3f s, GPU calls %.3f s, waiting for format+write %.3f s\n",
                          trace.acc[0], trace.acc[1], trace.acc[2]);
    trace.mark("statistics done");
    if (out_failed || !out.flush()) { fprintf(stderr, "ERROR: write to stdout failed\n"); return 1; }
    if (negative) { fprintf(stderr, "ERROR, cannot have negative coverage!!\n"); return 1; }
    fprintf(stderr, "STATS_GENERATION_TIME: %ld seconds.\n", (long)(time(NULL) - start_time));
    trace.mark("output complete");
    // Everything is written.  Leave without tearing gigabytes of tables, mappings and buffers down one by one: the
    // operating system and the driver reclaim them at process exit (this was 0.4 s of a 2.7 s run on a 20 M-read file).
    trace.wall("exit");
    fflush(stderr);
    _exit(0);
}
