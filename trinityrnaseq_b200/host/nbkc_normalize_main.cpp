// nbkc_normalize -- drop-in for util/support_scripts/nbkc_normalize.pl (the step right after the
// coverage statistics in util/insilico_read_normalization.pl:927,973).  Reads the (pair) statistics table, discards
// reads below --min_cov or with stdev/mean above --max_CV, and keeps the others with probability max_cov / median_cov
// using Perl's rand() after srand(12345) -- the same draws in the same order, so the selected accessions are the same
// bytes.  CPU only (the statistics were the GPU's job); one pass, no per-row heap traffic beyond the field strings.
#include <errno.h>

#include "perl_compat.hpp"

static const char* USAGE =
    "\n#############################################################\n#\n# Required:\n#\n"
    "#  --stats_file <string>     : pairs.stats.sorted\n#\n#  --max_cov <int>           : maximum coverage\n#\n"
    "#  --min_cov <int>           : minimum coverage\n#\n#  --max_CV <int>            : maximum coeff. var.\n#\n#\n"
    "#############################################################\n\n";

// accessions selected so far: Perl flushes STDOUT when it dies, so a malformed row late in the table must not swallow them
static std::string g_out;

static void fatal(const std::string& msg) {
    fwrite(g_out.data(), 1, g_out.size(), stdout);
    fflush(stdout);
    fprintf(stderr, "%s\n", msg.c_str());
    exit(255);
}

int main(int argc, char** argv) {
    perlc::LongOpts o(argc, argv, {"stats_file", "max_cov", "min_cov", "max_CV"}, {});
    const std::string stats_file = o.get("stats_file");
    const bool have_max_cov = o.has("max_cov") && perlc::LongOpts::is_int(o.get("max_cov"));
    const bool have_max_cv = o.has("max_CV") && perlc::LongOpts::is_int(o.get("max_CV"));
    long max_cov = have_max_cov ? atol(o.get("max_cov").c_str()) : 0;
    long min_cov = 1;
    if (o.has("min_cov") && perlc::LongOpts::is_int(o.get("min_cov"))) min_cov = atol(o.get("min_cov").c_str());
    const long max_cv = have_max_cv ? atol(o.get("max_CV").c_str()) : 0;
    if (stats_file.empty() || stats_file == "0" || max_cov == 0 || !have_max_cv) {      // Perl truthiness of the two scalars
        fputs(USAGE, stderr);
        return 255;
    }
    FILE* f = fopen(stats_file.c_str(), "r");
    if (!f) { fprintf(stderr, "%s at nbkc_normalize line 73.\n", strerror(errno ? errno : 2)); return errno ? errno : 2; }
    perlc::DelimReader rd(f, fatal);
    const int c_acc = rd.col("acc"), c_med = rd.col("median_cov"), c_sd = rd.col("stdev"), c_mean = rd.col("mean_cov");
    perlc::Drand48 rng(12345);
    unsigned long long total = 0, selected = 0, aberrant = 0, below = 0;
    std::vector<std::string> row;
    std::string& out = g_out;
    out.reserve(1 << 20);
    while (rd.next(row)) {
        total++;
        const double med = perlc::numify(perlc::DelimReader::field(row, c_med));
        if (med < (double)min_cov) { below++; continue; }
        const double sd = perlc::numify(perlc::DelimReader::field(row, c_sd));
        const double u = perlc::numify(perlc::DelimReader::field(row, c_mean));
        if (u <= 0) { aberrant++; continue; }
        const double cv = sd / u;
        if (cv > (double)max_cv) { aberrant++; continue; }
        if (med == 0) fatal("Illegal division by zero at nbkc_normalize line 111.");
        if (rng.next() <= (double)max_cov / med) {
            const std::string& acc = perlc::DelimReader::field(row, c_acc);
            size_t n = acc.size();
            if (n >= 2 && acc[n - 2] == '/' && (acc[n - 1] == '1' || acc[n - 1] == '2')) n -= 2;
            out.append(acc, 0, n);
            out.push_back('\n');
            selected++;
            if (out.size() > (1 << 20) - 4096) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    fclose(f);
    if (fflush(stdout) != 0) { fprintf(stderr, "write failed\n"); return 1; }
    if (!total) { fprintf(stderr, "Error, no reads made it to the normalization process...   at nbkc_normalize line 120.\n"); return 255; }
    fprintf(stderr, "%llu / %llu = %.2f%% reads selected during normalization.\n", selected, total, (double)selected / total * 100);
    fprintf(stderr, "%llu / %llu = %.2f%% reads discarded as likely aberrant based on coverage profiles.\n", aberrant, total,
            (double)aberrant / total * 100);
    fprintf(stderr, "%llu / %llu = %.2f%% reads discarded as below minimum coverage threshold=%ld\n", below, total,
            (double)below / total * 100, min_cov);
    return 0;
}
