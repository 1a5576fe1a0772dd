// jellyfish -- drop-in for the `jellyfish count | dump | histo | --version` calls Trinity makes
// (Trinity:2612-2632, 4065-4078; util/insilico_read_normalization.pl:617-643).  Jellyfish itself is a third-party
// binary (gmarcais/Jellyfish 2.3.0, Docker/Dockerfile:179) that is not part of the reference tree; this tool
// implements the same command lines and the same dump/histo text formats on top of libtrinity_gpu.
// The .jf file is a private binary format (only our own dump/histo read it, and Trinity deletes it afterwards).
#include <errno.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "fasta_io.hpp"
#include "seq_file.hpp"
#include "par_fasta.hpp"
#include "multi_gpu.hpp"
#include "tg_loader.hpp"
#include "tg_sidecar.hpp"

using namespace tgio;

namespace {

const char JF_MAGIC[8] = {'T', 'G', 'J', 'F', '0', '0', '1', '\n'};

struct JfHeader {
    char magic[8];
    uint32_t k;
    uint32_t canonical;
    uint64_t n;
    uint64_t histo[TG_HISTO_BINS];   // computed on the GPU at count time
};

[[noreturn]] void die_usage(const char* msg) {
    fprintf(stderr, "jellyfish: %s\n", msg);
    exit(1);
}

uint64_t parse_size(const std::string& s) {
    // "100000000", "100M", "2G", "1e8"
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) die_usage(("invalid size '" + s + "'").c_str());
    switch (*end) {
        case 'k': case 'K': v *= 1e3; break;
        case 'm': case 'M': v *= 1e6; break;
        case 'g': case 'G': v *= 1e9; break;
        case 't': case 'T': v *= 1e12; break;
        default: break;
    }
    return (uint64_t)v;
}

// getopt-free option scanner: supports "-m 25", "-m25", "--mer-len=25", "--mer-len 25"
struct Opts {
    std::vector<std::string> positional;
    std::vector<std::pair<std::string, std::string>> kv;
    bool has(const std::string& a, const std::string& b = "") const {
        for (auto& p : kv) if (p.first == a || (!b.empty() && p.first == b)) return true;
        return false;
    }
    std::string get(const std::string& a, const std::string& b, const std::string& def) const {
        for (auto& p : kv) if (p.first == a || p.first == b) return p.second;
        return def;
    }
};

Opts scan(int argc, char** argv, int first, const std::string& short_with_val, const std::vector<std::string>& long_with_val) {
    Opts o;
    for (int i = first; i < argc; i++) {
        std::string a = argv[i];
        if (a.size() >= 2 && a[0] == '-' && a[1] == '-') {
            size_t eq = a.find('=');
            std::string name = eq == std::string::npos ? a : a.substr(0, eq);
            bool takes = false;
            for (auto& l : long_with_val) if (l == name) takes = true;
            if (eq != std::string::npos) o.kv.push_back({name, a.substr(eq + 1)});
            else if (takes && i + 1 < argc) o.kv.push_back({name, argv[++i]});
            else o.kv.push_back({name, ""});
        } else if (a.size() >= 2 && a[0] == '-') {
            const char c = a[1];
            if (short_with_val.find(c) != std::string::npos) {
                if (a.size() > 2) o.kv.push_back({a.substr(0, 2), a.substr(2)});
                else if (i + 1 < argc) o.kv.push_back({a, argv[++i]});
                else die_usage(("option " + a + " needs a value").c_str());
            } else {
                for (size_t j = 1; j < a.size(); j++) o.kv.push_back({std::string("-") + a[j], ""});   // -Ct style clusters
            }
        } else {
            o.positional.push_back(a);
        }
    }
    return o;
}

// ---- sequence files -> record batches ----------------------------------------------------------------------
// FASTA (multi-line allowed: line breaks inside a record do not break k-mers) and FASTQ (4-line records).
// One GPU: chunks of parsed records are counted as they come.  Several GPUs (TRINITY_GPUS=0,1,..): every device has its
// own table and a worker thread with a one-chunk mailbox; the parser deals the chunks round robin, and the tables are
// summed into the first one at the end (export -> tg_table_load_pairs: counts add up, exactly like counting everything
// into one table).
struct CountWorker {
    tg_ctx* ctx = nullptr;
    tg_table* table = nullptr;
    int canonical = 0;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> box;
    bool full = false, stop = false;
    std::string error;
    void run() {
        for (;;) {
            std::vector<char> mine;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return full || stop; });
                if (!full) return;
                mine.swap(box);
                full = false;
            }
            cv.notify_all();
            if (error.empty() && tg_count_reads(table, mine.data(), mine.size(), canonical) != TG_OK) error = tg_last_error();
        }
    }
    void give(std::vector<char>& recs) {           // blocks while the previous chunk is still waiting to be taken
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !full; });
        box.swap(recs);
        full = true;
        lk.unlock();
        cv.notify_all();
    }
    void finish() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

void count_file(std::vector<CountWorker>& workers, size_t& next, tg_table* table, const std::string& path, int canonical) {
    FileView fv;
    std::string err;
    if (!fv.open(path, &err)) { fprintf(stderr, "jellyfish: %s\n", err.c_str()); exit(1); }
    std::vector<char> recs;
    // record bytes per counted chunk (TRINITY_GPU_COUNT_CHUNK: a test knob -- tiny chunks spread a tiny file over the devices)
    const bool knob = getenv("TRINITY_GPU_COUNT_CHUNK") != nullptr;
    const size_t FLUSH = knob ? (size_t)strtoull(getenv("TRINITY_GPU_COUNT_CHUNK"), nullptr, 10) : (256u << 20);
    auto consume = [&](std::vector<char>& r) {
        if (r.empty()) return;
        if (workers.empty()) {
            TGC(tg_count_reads(table, r.data(), r.size(), canonical));
        } else {
            workers[next % workers.size()].give(r);
            next++;
        }
        r.clear();
    };
    // FASTA (what Trinity feeds it: both.fa) splits at any line that starts with '>': chunks are parsed by a pool of threads
    // with the same parser and counted in file order (par_fasta.hpp) -- one thread walking a multi-gigabyte file was most of
    // the tool's wall time.  FASTQ records cannot be told from quality lines without context: one thread, as before.
    const char* q = fv.data;
    const char* fend = fv.data + fv.size;
    while (q < fend && (*q == '\n' || *q == '\r')) q++;
    const bool fasta = q < fend && *q == '>';
    const unsigned nthreads = host_threads(32);
    if (fasta && !knob && fv.size >= (64u << 20) && nthreads > 1) {
        OrderedChunkParser parser(fv.data, fv.size, 48u << 20, nthreads, nthreads + 2,
            [](const char* d, size_t n, RecordBatch& rb) {
                rb.recs.reserve(n);
                parse_sequence_file(d, d + n, rb.recs, (size_t)-1, [] {});
            });
        RecordBatch rb;
        while (parser.next(rb)) consume(rb.recs);
        return;
    }
    recs.reserve(FLUSH + (64u << 20));
    parse_sequence_file(fv.data, fv.data + fv.size, recs, FLUSH, [&]() {
        consume(recs);
        if (!workers.empty()) recs.reserve(FLUSH + (64u << 20));
    });
}

int cmd_count(int argc, char** argv) {
    Opts o = scan(argc, argv, 2, "mstocpLUQrFd", {"--mer-len", "--size", "--threads", "--output", "--counter-len",
                                                  "--out-counter-len", "--reprobes", "--lower-count", "--upper-count",
                                                  "--Files", "--generator", "--shell", "--bf-size", "--bc", "--if",
                                                  "--min-qual-char", "--quality-start", "--min-quality"});
    if (!o.has("-m", "--mer-len")) die_usage("count: missing required option -m, --mer-len");
    if (!o.has("-s", "--size")) die_usage("count: missing required option -s, --size");
    const int k = atoi(o.get("-m", "--mer-len", "0").c_str());
    if (k < 1 || k > 32) die_usage("count: mer length must be in 1..32 for the GPU k-mer table");
    const uint64_t size_hint = parse_size(o.get("-s", "--size", "0"));
    const std::string out_path = o.get("-o", "--output", "mer_counts.jf");
    const int canonical = o.has("-C", "--canonical") ? 1 : 0;
    std::vector<std::string> files = o.positional;
    if (files.empty()) files.push_back("/dev/fd/0");

    tgh::GpuSet gpus;
    gpus.open();
    tg_ctx* ctx = gpus.ctx[0];
    // -s is only an initial-size hint in jellyfish 2 (the hash grows); we additionally cap it by the input size so
    // that thousands of tiny phase-2 invocations do not each grab gigabytes
    uint64_t input_bytes = 0;
    for (auto& f : files) { struct stat st; if (stat(f.c_str(), &st) == 0 && S_ISREG(st.st_mode)) input_bytes += (uint64_t)st.st_size; else input_bytes += size_hint; }
    const uint64_t expected = std::min<uint64_t>(size_hint, input_bytes / 2 + 1024);
    tg_table* table = nullptr;
    TGC(tg_table_create(ctx, TG_TABLE_COUNT, k, expected, &table));
    std::vector<CountWorker> workers(gpus.size() > 1 ? gpus.size() : 0);
    for (size_t g = 0; g < workers.size(); g++) {
        workers[g].ctx = gpus.ctx[g];
        workers[g].canonical = canonical;
        if (g == 0) workers[g].table = table;
        else TGC(tg_table_create(gpus.ctx[g], TG_TABLE_COUNT, k, expected / gpus.size() + 1024, &workers[g].table));
        workers[g].th = std::thread([&workers, g] { workers[g].run(); });
    }
    size_t next = 0;
    for (auto& f : files) count_file(workers, next, table, f, canonical);
    for (auto& w : workers) w.finish();
    for (size_t g = 0; g < workers.size(); g++)
        if (!workers[g].error.empty()) { fprintf(stderr, "ERROR: GPU %zu: %s\n", g, workers[g].error.c_str()); return 3; }
    for (size_t g = 1; g < workers.size(); g++) {          // sum the other devices' tables into the first
        uint64_t* mk = nullptr; uint32_t* mc = nullptr; uint64_t mn = 0;
        TGC(tg_table_export(workers[g].table, 1, 0xFFFFFFFFu, /*sorted=*/0, 0, &mk, &mc, &mn));
        const uint64_t STEP = 32u << 20;
        for (uint64_t i = 0; i < mn; i += STEP) TGC(tg_table_load_pairs(table, mk + i, mc + i, std::min<uint64_t>(STEP, mn - i), canonical));
        tg_free(mk); tg_free(mc);
        tg_table_destroy(workers[g].table);
    }

    JfHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, JF_MAGIC, 8);
    h.k = (uint32_t)k; h.canonical = (uint32_t)canonical;
    TGC(tg_histo(table, h.histo));
    uint64_t* keys = nullptr; uint32_t* counts = nullptr; uint64_t n = 0;
    TGC(tg_table_export(table, 1, 0xFFFFFFFFu, /*sorted=*/1, canonical, &keys, &counts, &n));
    h.n = n;
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) { fprintf(stderr, "jellyfish: cannot open output file '%s': %s\n", out_path.c_str(), strerror(errno)); return 1; }
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    ok = ok && (n == 0 || fwrite(keys, 8, n, f) == n);
    ok = ok && (n == 0 || fwrite(counts, 4, n, f) == n);
    ok = (fclose(f) == 0) && ok;
    tg_free(keys); tg_free(counts);
    tg_table_destroy(table);
    gpus.close();
    if (!ok) { fprintf(stderr, "jellyfish: write to '%s' failed\n", out_path.c_str()); unlink(out_path.c_str()); return 1; }
    return 0;
}

struct JfFile {
    FileView fv;
    const JfHeader* h = nullptr;
    const uint64_t* keys = nullptr;
    const uint32_t* counts = nullptr;
    void open(const std::string& path) {
        std::string err;
        if (!fv.open(path, &err)) { fprintf(stderr, "jellyfish: %s\n", err.c_str()); exit(1); }
        if (fv.size < sizeof(JfHeader) || memcmp(fv.data, JF_MAGIC, 8) != 0) {
            fprintf(stderr, "jellyfish: '%s' is not a k-mer database written by this jellyfish\n", path.c_str());
            exit(1);
        }
        h = (const JfHeader*)fv.data;
        if (fv.size < sizeof(JfHeader) + h->n * 12) { fprintf(stderr, "jellyfish: '%s' is truncated\n", path.c_str()); exit(1); }
        keys = (const uint64_t*)(fv.data + sizeof(JfHeader));
        counts = (const uint32_t*)(fv.data + sizeof(JfHeader) + h->n * 8);
    }
};

int cmd_dump(int argc, char** argv) {
    Opts o = scan(argc, argv, 2, "LUo", {"--lower-count", "--upper-count", "--output"});
    if (o.positional.size() != 1) die_usage("dump: exactly one database argument expected");
    const uint64_t lower = o.has("-L", "--lower-count") ? strtoull(o.get("-L", "--lower-count", "0").c_str(), nullptr, 10) : 0;
    const uint64_t upper = o.has("-U", "--upper-count") ? strtoull(o.get("-U", "--upper-count", "0").c_str(), nullptr, 10) : ~0ull;
    const bool column = o.has("-c", "--column"), tab = o.has("-t", "--tab");
    JfFile jf;
    jf.open(o.positional[0]);
    int fd = 1;
    if (o.has("-o", "--output")) {
        fd = ::open(o.get("-o", "--output", "").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) { fprintf(stderr, "jellyfish: cannot open output file: %s\n", strerror(errno)); return 1; }
    }
    const int k = (int)jf.h->k;
    // Binary hand-off (tg_sidecar.hpp): a FASTA dump that lands in a regular file gets `<file>.tgk` next to it with the
    // same records as packed pairs, so fastaToKmerCoverageStats --kmers need not re-parse ~31 B of text per k-mer.
    const std::string text_path = (!column && !tgside::disabled()) ? tgside::regular_file_behind(fd) : "";
    // The sidecar is streamed like the text (O(1) memory, whatever the table size): header placeholder, the packed k-mers as
    // they are printed, then -- in a second pass over the memory-mapped counts with the same filter -- the counts, and last the
    // real header.  Best effort: a sidecar that cannot be written is simply absent (the FASTA is complete either way).
    const std::string side = text_path + ".tgk", side_tmp = side + ".tmp";
    FILE* sf = text_path.empty() ? nullptr : fopen(side_tmp.c_str(), "wb");
    bool side_ok = sf != nullptr;
    uint64_t side_n = 0;
    tgside::TgkHeader sh;
    memset(&sh, 0, sizeof sh);
    if (sf) {
        setvbuf(sf, nullptr, _IOFBF, 4u << 20);
        side_ok = fwrite(&sh, sizeof sh, 1, sf) == 1;
    }
    tgside::TextHash hash;
    {
        // The records are formatted by a pool of threads, a block of BLOCK k-mers at a time (thread w takes the w-th slice
        // of the block and writes its lines -- and, for the sidecar, the packed k-mers it printed -- into buffers of its
        // own); the main thread writes block i, hashes it and appends the sidecar keys while block i + 1 is being formatted.
        // Same bytes in the same order as one loop over the table; memory stays O(BLOCK).
        auto put_uint = [](char* dst, uint32_t v) {             // decimal without sprintf: this loop runs 10^8 times
            char tmp[12]; int n = 0;
            do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
            for (int i = 0; i < n; i++) dst[i] = tmp[n - 1 - i];
            return n;
        };
        const uint64_t total = jf.h->n;
        const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(host_threads(32), total / 65536 + 1));
        const uint64_t BLOCK = 1u << 21;
        struct Block { std::vector<std::vector<char>> text; std::vector<std::vector<uint64_t>> keys; };
        Block blocks[2];
        for (auto& b : blocks) { b.text.resize(nt); b.keys.resize(nt); }
        const bool want_keys = sf != nullptr;
        auto format_block = [&](Block& blk, uint64_t b0) {
            const uint64_t b1 = std::min(total, b0 + BLOCK);
            parallel_for_threads(nt, [&](unsigned w) {
                std::vector<char>& buf = blk.text[w];
                std::vector<uint64_t>& kk = blk.keys[w];
                buf.clear(); kk.clear();
                const uint64_t a = b0 + (b1 - b0) * w / nt, e = b0 + (b1 - b0) * (w + 1) / nt;
                buf.resize((size_t)(e - a) * (size_t)(k + 14));          // (longest record: '>' + 10 digits + 2 newlines + k)
                char* line = buf.data();
                for (uint64_t i = a; i < e; i++) {
                    const uint32_t c = jf.counts[i];
                    if (c < lower || c > upper) continue;
                    int n;
                    if (column) {                                       // "KMER COUNT"
                        tgh::unpack_kmer(jf.keys[i], k, line);
                        n = k;
                        line[n++] = tab ? '\t' : ' ';
                        n += put_uint(line + n, c);
                        line[n++] = '\n';
                    } else {                                            // FASTA: ">COUNT\nKMER\n"
                        line[0] = '>';
                        n = 1 + put_uint(line + 1, c);
                        line[n++] = '\n';
                        tgh::unpack_kmer(jf.keys[i], k, line + n);
                        n += k;
                        line[n++] = '\n';
                    }
                    line += n;
                    if (want_keys) kk.push_back(jf.keys[i]);
                }
                buf.resize((size_t)(line - buf.data()));
            });
        };
        bool write_ok = true;
        auto write_all = [&](const char* p, size_t n) {               // (the slices are megabytes: no staging copy)
            while (n && write_ok) {
                const ssize_t w = ::write(fd, p, n);
                if (w < 0) { if (errno == EINTR) continue; write_ok = false; break; }
                p += w; n -= (size_t)w;
            }
        };
        std::thread ahead;
        if (total) format_block(blocks[0], 0);
        for (uint64_t b0 = 0, it = 0; b0 < total; b0 += BLOCK, it++) {
            Block& cur = blocks[it & 1];
            if (b0 + BLOCK < total) ahead = std::thread([&, it, b0] { format_block(blocks[(it + 1) & 1], b0 + BLOCK); });
            std::thread side_thread;                      // the sidecar's share of the block (hash of the text, packed k-mers) beside the write
            if (sf) side_thread = std::thread([&] {
                for (unsigned w = 0; w < nt; w++) {
                    hash.update(cur.text[w].data(), cur.text[w].size());
                    if (!cur.keys[w].empty())
                        side_ok = side_ok && fwrite(cur.keys[w].data(), 8, cur.keys[w].size(), sf) == cur.keys[w].size();
                    side_n += cur.keys[w].size();
                }
            });
            for (unsigned w = 0; w < nt; w++) write_all(cur.text[w].data(), cur.text[w].size());
            if (side_thread.joinable()) side_thread.join();
            if (ahead.joinable()) ahead.join();
        }
        if (!write_ok) { fprintf(stderr, "jellyfish: write failed: %s\n", strerror(errno)); if (sf) { fclose(sf); unlink(side_tmp.c_str()); } return 1; }
    }
    if (fd != 1) ::close(fd);
    if (sf) {
        std::vector<uint32_t> pass;                       // the counts that passed the filter, a million at a time
        pass.reserve(1u << 20);
        for (uint64_t i = 0; side_ok && i < jf.h->n; i++) {
            const uint32_t c = jf.counts[i];
            if (c < lower || c > upper) continue;
            pass.push_back(c);
            if (pass.size() == (1u << 20)) { side_ok = fwrite(pass.data(), 4, pass.size(), sf) == pass.size(); pass.clear(); }
        }
        if (side_ok && !pass.empty()) side_ok = fwrite(pass.data(), 4, pass.size(), sf) == pass.size();
        memcpy(sh.magic, tgside::TGK_MAGIC, 8);
        sh.k = (uint32_t)k; sh.n = side_n; sh.text_bytes = hash.total; sh.text_hash = hash.digest();
        side_ok = side_ok && fseek(sf, 0, SEEK_SET) == 0 && fwrite(&sh, sizeof sh, 1, sf) == 1;
        side_ok = (fclose(sf) == 0) && side_ok;
        if (side_ok) side_ok = rename(side_tmp.c_str(), side.c_str()) == 0;
        if (!side_ok) unlink(side_tmp.c_str());
    }
    return 0;
}

int cmd_histo(int argc, char** argv) {
    Opts o = scan(argc, argv, 2, "lhito", {"--low", "--high", "--increment", "--threads", "--output"});
    if (o.positional.size() != 1) die_usage("histo: exactly one database argument expected");
    const uint64_t low = strtoull(o.get("-l", "--low", "1").c_str(), nullptr, 10);
    const uint64_t high = strtoull(o.get("-h", "--high", "10000").c_str(), nullptr, 10);
    const uint64_t inc = strtoull(o.get("-i", "--increment", "1").c_str(), nullptr, 10);
    const bool full = o.has("-f", "--full");
    if (inc == 0) die_usage("histo: increment must be positive");
    if (high > 10000) die_usage("histo: --high above 10000 is not supported (bins are accumulated on the GPU up to 10000)");
    // counts above 10000 are one bin here; that is exact only while they all fall into the LAST bucket, i.e. while the
    // ceiling high + increment does not pass 10001 (the defaults -h 10000 -i 1 are exactly that)
    if (high + inc > 10001) die_usage("histo: --high + --increment above 10001 is not supported (counts above 10000 are one bin)");
    JfFile jf;
    jf.open(o.positional[0]);
    // jellyfish 2 histo_main: base = low>0 ? (inc>=low ? 0 : low-inc) : 0; ceil = high+inc; bucket = (val-base)/inc,
    // below base -> first bucket, above ceil -> last bucket; buckets printed as "base+i*inc count"
    const uint64_t base = low > 0 ? (inc >= low ? 0 : low - inc) : 0;
    const uint64_t ceil = high + inc;
    const uint64_t nb = (ceil + inc - base) / inc;
    std::vector<uint64_t> hist(nb, 0);
    for (uint64_t c = 0; c < TG_HISTO_BINS; c++) {
        const uint64_t m = jf.h->histo[c];
        if (!m) continue;
        const bool overflow = c == TG_HISTO_BINS - 1;     // every count > 10000
        uint64_t b;
        if (c < base) b = 0;
        else if (overflow || c > ceil) b = nb - 1;
        else b = (c - base) / inc;
        if (b >= nb) b = nb - 1;
        hist[b] += m;
    }
    FILE* f = stdout;
    if (o.has("-o", "--output")) {
        f = fopen(o.get("-o", "--output", "").c_str(), "w");
        if (!f) { fprintf(stderr, "jellyfish: cannot open output file: %s\n", strerror(errno)); return 1; }
    }
    for (uint64_t i = 0; i < nb; i++)
        if (hist[i] > 0 || full) fprintf(f, "%llu %llu\n", (unsigned long long)(base + i * inc), (unsigned long long)hist[i]);
    if (f != stdout && fclose(f) != 0) { fprintf(stderr, "jellyfish: write failed\n"); return 1; }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 2 && (!strcmp(argv[1], "--version") || !strcmp(argv[1], "-V"))) {
        printf("jellyfish 2.3.0\n");      // Trinity requires the second token to match /^2\./ (Trinity:4070-4074)
        return 0;
    }
    if (argc < 2) die_usage("usage: jellyfish <count|dump|histo|--version> [options]");
    const std::string cmd = argv[1];
    if (cmd == "count") return cmd_count(argc, argv);
    if (cmd == "dump") return cmd_dump(argc, argv);
    if (cmd == "histo") return cmd_histo(argc, argv);
    die_usage(("unknown subcommand '" + cmd + "' (supported: count, dump, histo)").c_str());
}
