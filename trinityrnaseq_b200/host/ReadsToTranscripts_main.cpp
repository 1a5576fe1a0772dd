// ReadsToTranscripts -- drop-in for Chrysalis/bin/ReadsToTranscripts (reference:
// Chrysalis/analysis/ReadsToTranscripts.cc).  Same argv, same output files (-o and -o.rcts.out), same exit codes;
// bundle k-mer labelling and the per-read vote run on the GPU through libtrinity_gpu.
//
// Output order: the reference emits, per chunk of -max_mem_reads reads, the assigned reads grouped by bundle index
// ascending and (single-threaded) by read order inside a bundle (multimap insertion order,
// ReadsToTranscripts.cc:276-343).  We reproduce exactly that `-t 1` order, which is deterministic.
#include <errno.h>

#include <algorithm>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "fasta_io.hpp"
#include "par_fasta.hpp"
#include "multi_gpu.hpp"
#include "tg_loader.hpp"

using namespace tgio;

namespace {

const size_t PART_BYTES = 48u << 20;          // FASTA bytes per parsed part
const size_t LINES_PER_BLOCK = 1u << 21;      // output lines formatted (and held in memory) at a time

struct ArgDef { const char* name; const char* desc; bool is_bool; const char* def; };
const ArgDef ARGS[] = {
    {"-i", "reads fasta", false, nullptr},
    {"-f", "fasta input file (concatenated flat components)", false, nullptr},
    {"-o", "output file", false, nullptr},
    {"-max_mem_reads", "Maximum number of reads to load into memory", false, "-1"},
    {"-strand", "strand specific data", true, "0"},
    {"-p", "percent of read kmers require mapping to component", false, "0"},
    {"-verbose", "prints more status info", true, "0"},
    {"-t", "number of threads (default: env OMP_NUM_THREADS)", false, "0"},
    {"-min_kmer_entropy", "min kmer entropy for assigning reads to iworm contigs", false, "1.5"},
};

void show_help(const char* argv0) {
    printf("\n%s: Assigns reads to graph components.\n\n\nAvailable arguments:\n", argv0);
    for (const ArgDef& a : ARGS) {
        printf("\n%s<%s> : %s", a.name, a.is_bool ? "bool" : "string", a.desc);
        if (a.def) printf(", default=%s", a.def);
    }
    printf("\n\n");
}

bool is_float_token(const std::string& s) {
    for (char c : s) if ((c < '0' || c > '9') && c != '.') return false;
    return true;
}

// commandLineParser::parse (Chrysalis/base/CommandLineParser.h:169-273): registered names only; the next token is
// the value unless it starts with '-' (then the flag gets the empty value and only one token is consumed)
std::map<std::string, std::string> parse_args(int argc, char** argv) {
    if (argc == 1) { show_help(argv[0]); exit(-1); }
    for (int i = 1; i < argc; i++) if (std::string(argv[i]) == "-h") { show_help(argv[0]); exit(-1); }
    std::set<std::string> names;
    for (const ArgDef& a : ARGS) names.insert(a.name);
    std::map<std::string, std::string> nv;
    int i = 1;
    while (i < argc) {
        std::string n(argv[i]), v = i < argc - 1 ? argv[i + 1] : "";
        if (!names.count(n)) {
            if (n == "-print-command-line") {
                printf("----------------------- Welcome to Trinity -------------------------------\nThis module was invoked via:\n");
                for (int k = 0; k < argc; k++) printf("%s ", argv[k]);
                printf("\n----------------------- Welcome to Trinity -------------------------------\n\n");
                if (argc == 2) { show_help(argv[0]); exit(-1); }
                i++;
                continue;
            }
            printf("\nInvalid command-line arg: %s\n", n.c_str());
            show_help(argv[0]);
            exit(-1);
        }
        // second whitespace token of the value being numeric lets a negative-looking value through (is_float rule)
        bool is_float = false;
        {
            size_t sp = v.find_first_of(" \t");
            if (sp != std::string::npos) {
                size_t b = v.find_first_not_of(" \t", sp);
                if (b != std::string::npos) {
                    size_t e = v.find_first_of(" \t", b);
                    is_float = is_float_token(v.substr(b, e == std::string::npos ? e : e - b));
                }
            }
        }
        const bool v_dash = !v.empty() && v[0] == '-';
        if (is_float || !v_dash) { nv.insert({n, v}); i += 2; }
        else { nv.insert({n, ""}); i += 1; }
    }
    return nv;
}

std::string get_str(const std::map<std::string, std::string>& nv, const char* key) {
    auto it = nv.find(key);
    if (it == nv.end()) { printf("need to specify %s\n", key); exit(-1); }
    return it->second;
}
std::string get_opt(const std::map<std::string, std::string>& nv, const char* key, const char* def) {
    auto it = nv.find(key);
    return (it == nv.end() || it->second.empty()) ? std::string(def) : it->second;
}

}  // namespace

int main(int argc, char** argv) {
    auto nv = parse_args(argc, argv);
    fprintf(stderr, "-------------------------------------------\n---- Chrysalis: ReadsToTranscripts --------\n"
                    "-- (Place reads on Inchworm Bundles) ------\n-------------------------------------------\n\n");
    const std::string reads_file = get_str(nv, "-i");
    const std::string out_file = get_str(nv, "-o");
    const std::string bundle_file = get_str(nv, "-f");
    long max_mem_reads = (int)atol(get_opt(nv, "-max_mem_reads", "-1").c_str());   // GetLongValueFor returns int
    const bool strand = nv.count("-strand") != 0;
    const int pct_required = atoi(get_opt(nv, "-p", "0").c_str());
    const int num_threads = atoi(get_opt(nv, "-t", "0").c_str());
    const bool verbose = nv.count("-verbose") != 0;
    const float min_kmer_entropy = (float)atof(get_opt(nv, "-min_kmer_entropy", "1.5").c_str());
    if (max_mem_reads > 0) fprintf(stderr, "Setting maximum number of reads to load in memory to %ld\n", max_mem_reads);
    else max_mem_reads = 2147483647;
    if (num_threads > 0) fprintf(stderr, "-setting num threads to: %d\n", num_threads);   // accepted; the work is on the GPU
    const int k = 25;

    tgh::GpuSet gpus;                       // TRINITY_GPUS=0,1,..: every device labels its own copy of the (small) bundle
    gpus.open();                            // table, the reads of a chunk are split over the devices (multi_gpu.hpp)
    std::string err;

    // ---- bundles ----------------------------------------------------------------------------------------
    fprintf(stderr, "Reading bundled inchworm contigs... \n");
    RecordBatch bundles;
    std::vector<std::string> bundle_names;
    {
        FileView bf;
        if (!bf.open(bundle_file, &err)) { fprintf(stderr, "Could not open file for read: %s\n", bundle_file.c_str()); return 1; }
        read_bundles(bf.data, bf.size, bundles, bundle_names);
    }
    fprintf(stderr, "done!\n");
    // component number = atoi(name.substr(3)) (strip ">s_"), ReadsToTranscripts.cc:153-156
    std::vector<int> component_no(bundle_names.size());
    for (size_t i = 0; i < bundle_names.size(); i++)
        component_no[i] = bundle_names[i].size() > 3 ? atoi(bundle_names[i].c_str() + 3) : 0;

    std::vector<tg_table*> tables(gpus.size(), nullptr);
    fprintf(stderr, "Assigning kmers to Iworm bundles ... ");
    tgh::on_every_gpu(gpus.size(), [&](size_t g) -> int {
        int rc = tg_table_create(gpus.ctx[g], TG_TABLE_LABEL, k, bundles.recs.size() + 1024, &tables[g]);
        if (rc == TG_OK) rc = tg_label_bundles(tables[g], bundles.recs.data(), bundles.offs.data(), bundles.count(), 0);
        return rc;
    });
    fprintf(stderr, "done!\n");
    std::vector<uint8_t> entropy_ok(26 * 26 * 26);
    tg_entropy_table(k, min_kmer_entropy, entropy_ok.data());

    // ---- reads, in chunks of max_mem_reads ------------------------------------------------------------------
    // The file is cut at header lines into parts that a pool of threads parses with the reference reader's rules
    // (par_fasta.hpp; in file order).  A part is assigned on the GPU as soon as it arrives -- the pool is already parsing
    // the next ones -- and a chunk of max_mem_reads reads is simply a run of (pieces of) parts: grouped by bundle over the
    // whole chunk and written like the reference writes it, its lines formatted block by block on the pool's cores.
    FileView rf;
    if (!rf.open(reads_file, &err)) { fprintf(stderr, "ERROR: %s\n", err.c_str()); return 1; }
    int out_fd = ::open(out_file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (out_fd < 0) { fprintf(stderr, "error writing file %s: %s\n", out_file.c_str(), strerror(errno)); return 1; }
    fprintf(stderr, "Processing reads:\n");
    const unsigned nthreads_host = host_threads(32);
    OrderedChunkParser parser(rf.data, rf.size, PART_BYTES, nthreads_host, nthreads_host + 2,
        [](const char* d, size_t n, RecordBatch& rb) {
            rb.recs.reserve(n); rb.offs.reserve(n / 48 + 16); rb.names.reserve(n / 4); rb.name_offs.reserve(n / 48 + 16);
            DnaStreamReader rd(d, n);
            const char* name; size_t name_len;
            while (rd.next(&name, &name_len, rb.recs)) {          // sequence lands in the batch directly
                rb.end_record();
                rb.add_name(name, name_len);
            }
        }, /*after_sequence_line=*/true);

    struct Piece { std::shared_ptr<RecordBatch> rb; size_t a, b, base; };      // reads [a, b) of a part = reads [base, ..) of the chunk
    unsigned long read_count = 0, total_reads_read = 0;
    std::vector<int32_t> best, pct;
    std::vector<uint32_t> order, bucket_start;
    const size_t nb = bundle_names.size();
    // host memory guard: the reference keeps a whole chunk in RAM too, but 2^31 reads is its "unlimited"
    const size_t CHUNK_BYTES_SOFT = (size_t)8 << 30;
    std::shared_ptr<RecordBatch> carry;          // a part whose tail belongs to the next chunk
    size_t carry_pos = 0;
    bool more = true, write_ok = true;
    auto write_all = [&](const char* p, size_t n) {
        while (n && write_ok) {
            const ssize_t w = ::write(out_fd, p, n);
            if (w < 0) { if (errno == EINTR) continue; write_ok = false; break; }
            p += w; n -= (size_t)w;
        }
    };
    while (more) {
        fprintf(stderr, " reading another %ld... ", max_mem_reads);
        std::vector<Piece> pieces;
        size_t got = 0, got_bytes = 0;
        best.clear(); pct.clear();
        while ((long)got < max_mem_reads) {
            std::shared_ptr<RecordBatch> part = carry;
            size_t a = carry_pos;
            carry.reset(); carry_pos = 0;
            if (!part) {
                part = std::make_shared<RecordBatch>();
                if (!parser.next(*part)) { more = false; break; }
                a = 0;
            }
            const size_t avail = part->count() - a;
            if (avail == 0) continue;
            const size_t take = std::min<size_t>(avail, (size_t)max_mem_reads - got);
            if (take < avail) { carry = part; carry_pos = a + take; }
            best.resize(got + take); pct.resize(got + take);
            {   // the vote of these reads, now: the pool is parsing the parts behind this one meanwhile
                const uint64_t* offs = part->offs.data() + a;
                const auto ranges = tgh::split_reads_by_bytes(offs, take, gpus.size());
                tgh::on_every_gpu(gpus.size(), [&](size_t g) -> int {
                    const size_t x = ranges[g].first, y = ranges[g].second;
                    if (x == y) return TG_OK;
                    return tg_assign_reads(tables[g], part->recs.data(), offs + x, y - x, strand, entropy_ok.data(),
                                           best.data() + got + x, pct.data() + got + x, nullptr);
                });
            }
            pieces.push_back(Piece{part, a, a + take, got});
            got += take;
            got_bytes += (size_t)(part->offs[a + take] - part->offs[a]);
            if (max_mem_reads == 2147483647 && got_bytes > CHUNK_BYTES_SOFT) break;   // "unlimited": bound RAM
        }
        if (got == 0) { fprintf(stderr, "finished reading reads\n"); break; }
        fprintf(stderr, "done.  Read %zu reads.\n", got);
        const size_t n = got;
        total_reads_read += n;
        fprintf(stderr, "[%lu] reads analyzed for mapping.\n", total_reads_read);
        // read i of the chunk -> its piece (pieces are few: a binary search per use)
        auto piece_of = [&](size_t i) -> const Piece& {
            size_t lo = 0, hi = pieces.size();
            while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (pieces[mid].base <= i) lo = mid; else hi = mid; }
            return pieces[lo];
        };

        // group by bundle index ascending, read order inside a bundle (stable counting sort)
        bucket_start.assign(nb + 1, 0);
        size_t assigned = 0;
        for (size_t i = 0; i < n; i++) {
            const bool ok = best[i] != -1 && pct[i] >= pct_required;
            if (ok) { bucket_start[(size_t)best[i] + 1]++; assigned++; }
            else {
                best[i] = -1;
                if (verbose) {
                    const Piece& pc = piece_of(i);
                    const size_t r = pc.a + (i - pc.base);
                    fprintf(stderr, "WARNING: No component mapping for read: %.*s : %.*s\n", (int)pc.rb->name_len(r), pc.rb->name(r),
                            (int)pc.rb->seq_len(r), pc.rb->seq(r));
                }
            }
        }
        size_t components_written = 0;
        for (size_t b = 0; b < nb; b++) { if (bucket_start[b + 1]) components_written++; bucket_start[b + 1] += bucket_start[b]; }
        order.resize(assigned);
        {
            std::vector<uint32_t> cur(bucket_start.begin(), bucket_start.end() - 1);
            for (size_t i = 0; i < n; i++) if (best[i] >= 0) order[cur[(size_t)best[i]]++] = (uint32_t)i;
        }
        // the lines, LINES_PER_BLOCK at a time: thread w formats the w-th slice of the block into a buffer of its own, the
        // main thread writes block j while block j + 1 is being formatted
        const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(nthreads_host, assigned / 4096 + 1));
        std::vector<std::vector<char>> bufs[2];
        bufs[0].resize(nt); bufs[1].resize(nt);
        auto format_block = [&](std::vector<std::vector<char>>& out_bufs, size_t j0) {
            const size_t j1 = std::min(assigned, j0 + LINES_PER_BLOCK);
            parallel_for_threads(nt, [&](unsigned w) {
                std::vector<char>& buf = out_bufs[w];
                buf.clear();
                std::string nm;
                char num[24];
                auto put_int = [&](long long v) {
                    int m = 0; bool neg = v < 0; unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
                    do { num[m++] = (char)('0' + u % 10); u /= 10; } while (u);
                    if (neg) buf.push_back('-');
                    while (m) buf.push_back(num[--m]);
                };
                const size_t a = j0 + (j1 - j0) * w / nt, e = j0 + (j1 - j0) * (w + 1) / nt;
                for (size_t j = a; j < e; j++) {
                    const size_t i = order[j];
                    const Piece& pc = piece_of(i);
                    const size_t r = pc.a + (i - pc.base);
                    put_int(component_no[(size_t)best[i]]);
                    buf.push_back('\t');
                    format_read_name(pc.rb->name(r), pc.rb->name_len(r), nm);
                    buf.insert(buf.end(), nm.begin(), nm.end());
                    buf.push_back('\t');
                    put_int(pct[i]);
                    buf.push_back('%'); buf.push_back('\t');
                    buf.insert(buf.end(), pc.rb->seq(r), pc.rb->seq(r) + pc.rb->seq_len(r));          // original case, as read
                    buf.push_back('\n');
                }
            });
        };
        if (assigned) format_block(bufs[0], 0);
        for (size_t j0 = 0, it = 0; j0 < assigned; j0 += LINES_PER_BLOCK, it++) {
            std::thread ahead;
            if (j0 + LINES_PER_BLOCK < assigned) ahead = std::thread([&, it, j0] { format_block(bufs[(it + 1) & 1], j0 + LINES_PER_BLOCK); });
            for (unsigned w = 0; w < nt; w++) write_all(bufs[it & 1][w].data(), bufs[it & 1][w].size());
            if (ahead.joinable()) ahead.join();
        }
        read_count += assigned;
        if (!write_ok) { fprintf(stderr, "error writing file %s: %s\n", out_file.c_str(), strerror(errno)); return 1; }
        if (components_written) fprintf(stderr, "[%zu] components written.\n", components_written);
    }
    ::close(out_fd);
    fprintf(stderr, "Done\n");

    const std::string rc_file = out_file + ".rcts.out";
    FILE* f = fopen(rc_file.c_str(), "w");
    if (!f) { fprintf(stderr, "cannot write %s\n", rc_file.c_str()); return 1; }
    fprintf(f, "%lu\n", read_count);
    fclose(f);
    for (tg_table* t : tables) tg_table_destroy(t);
    gpus.close();
    return 0;
}
