// ReadsToTranscripts -- drop-in for Chrysalis/bin/ReadsToTranscripts (reference:
// Chrysalis/analysis/ReadsToTranscripts.cc).  Same argv, same output files (-o and -o.rcts.out), same exit codes;
// bundle k-mer labelling and the per-read vote run on the GPU through libtrinity_gpu.
//
// Output order: the reference emits, per chunk of -max_mem_reads reads, the assigned reads grouped by bundle index
// ascending and (single-threaded) by read order inside a bundle (multimap insertion order,
// ReadsToTranscripts.cc:276-343).  We reproduce exactly that `-t 1` order, which is deterministic.
#include <errno.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "fasta_io.hpp"
#include "multi_gpu.hpp"
#include "tg_loader.hpp"

using namespace tgio;

namespace {

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
    FileView rf;
    if (!rf.open(reads_file, &err)) { fprintf(stderr, "ERROR: %s\n", err.c_str()); return 1; }
    int out_fd = ::open(out_file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (out_fd < 0) { fprintf(stderr, "error writing file %s: %s\n", out_file.c_str(), strerror(errno)); return 1; }
    OutBuf out(out_fd, 32u << 20);
    DnaStreamReader rd(rf.data, rf.size);
    fprintf(stderr, "Processing reads:\n");

    unsigned long read_count = 0, total_reads_read = 0;
    RecordBatch rb;
    std::vector<int32_t> best, pct;
    std::vector<uint32_t> order, bucket_start;
    std::string nm;
    const size_t nb = bundle_names.size();
    // host memory guard: the reference keeps a whole chunk in RAM too, but 2^31 reads is its "unlimited"
    const size_t CHUNK_BYTES_SOFT = (size_t)8 << 30;
    bool more = true;
    while (more) {
        fprintf(stderr, " reading another %ld... ", max_mem_reads);
        rb.clear();
        long got = 0;
        const char* name; size_t name_len;
        while (got < max_mem_reads) {
            if (!rd.next(&name, &name_len, rb.recs)) { more = false; break; }   // sequence lands in the batch directly
            rb.end_record();
            rb.add_name(name, name_len);
            got++;
            if (max_mem_reads == 2147483647 && rb.recs.size() > CHUNK_BYTES_SOFT) break;   // "unlimited": bound RAM
        }
        if (got == 0) { fprintf(stderr, "finished reading reads\n"); break; }
        fprintf(stderr, "done.  Read %ld reads.\n", got);
        const size_t n = rb.count();
        best.resize(n); pct.resize(n);
        {
            const auto ranges = tgh::split_reads_by_bytes(rb.offs.data(), n, gpus.size());
            tgh::on_every_gpu(gpus.size(), [&](size_t g) -> int {
                const size_t a = ranges[g].first, b = ranges[g].second;
                if (a == b) return TG_OK;
                return tg_assign_reads(tables[g], rb.recs.data(), rb.offs.data() + a, b - a, strand, entropy_ok.data(),
                                       best.data() + a, pct.data() + a, nullptr);
            });
        }
        total_reads_read += n;
        fprintf(stderr, "[%lu] reads analyzed for mapping.\n", total_reads_read);

        // group by bundle index ascending, read order inside a bundle (stable counting sort)
        bucket_start.assign(nb + 1, 0);
        size_t assigned = 0;
        for (size_t i = 0; i < n; i++) {
            const bool ok = best[i] != -1 && pct[i] >= pct_required;
            if (ok) { bucket_start[(size_t)best[i] + 1]++; assigned++; }
            else { best[i] = -1; if (verbose) fprintf(stderr, "WARNING: No component mapping for read: %.*s : %.*s\n",
                                                     (int)rb.name_len(i), rb.name(i), (int)rb.seq_len(i), rb.seq(i)); }
        }
        size_t components_written = 0;
        for (size_t b = 0; b < nb; b++) { if (bucket_start[b + 1]) components_written++; bucket_start[b + 1] += bucket_start[b]; }
        order.resize(assigned);
        {
            std::vector<uint32_t> cur(bucket_start.begin(), bucket_start.end() - 1);
            for (size_t i = 0; i < n; i++) if (best[i] >= 0) order[cur[(size_t)best[i]]++] = (uint32_t)i;
        }
        for (size_t j = 0; j < assigned; j++) {
            const size_t i = order[j];
            out.put_int(component_no[(size_t)best[i]]);
            out.putc('\t');
            format_read_name(rb.name(i), rb.name_len(i), nm);
            out.put(nm);
            out.putc('\t');
            out.put_int(pct[i]);
            out.put("%\t", 2);
            out.put(rb.seq(i), rb.seq_len(i));          // original case, as read
            out.putc('\n');
        }
        read_count += assigned;
        if (out.failed()) { fprintf(stderr, "error writing file %s: %s\n", out_file.c_str(), strerror(errno)); return 1; }
        if (components_written) fprintf(stderr, "[%zu] components written.\n", components_written);
    }
    if (!out.flush()) { fprintf(stderr, "error writing file %s: %s\n", out_file.c_str(), strerror(errno)); return 1; }
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
