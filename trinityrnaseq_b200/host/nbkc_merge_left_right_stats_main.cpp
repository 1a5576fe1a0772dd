// nbkc_merge_left_right_stats -- drop-in for util/support_scripts/nbkc_merge_left_right_stats.pl (SURVEY §8f rank 3;
// util/insilico_read_normalization.pl:959): joins the sorted per-read statistics of the left and right reads into one
// line per pair whose median / mean / stdev are the averages, printed "%.1f" like the script.  CPU only.
#include <errno.h>

#include "perl_compat.hpp"

static const char* USAGE =
    "\n###############################################################################\n#\n# Required:\n#\n"
    "#  --left <string>     left.fq.stats \n#  --right <string>    right.fq.stats\n#\n# Optional\n#\n"
    "#  --sorted            flag indicating that entries are lexically sorted\n"
    "#                       (this can account for differences in representation by \n"
    "#                        reads in either file)\n#                       Unpaired entries are ignored.\n#\n"
    "################################################################################\n\n\n";

static void fatal(const std::string& msg) {
    fprintf(stderr, "%s\n", msg.c_str());
    exit(255);
}

static bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static FILE* open_stats(const std::string& path) {
    FILE* f;
    if (ends_with(path, ".gz")) f = popen(("gunzip -c " + path).c_str(), "r");
    else if (ends_with(path, ".xz")) f = popen(("xz -cd " + path).c_str(), "r");
    else f = fopen(path.c_str(), "r");
    if (!f) { fprintf(stderr, "%s at nbkc_merge_left_right_stats line 60.\n", strerror(errno ? errno : 2)); exit(errno ? errno : 2); }
    return f;
}

// $acc =~ /^(\S+)\/\d$/ ? $1 : $acc
static std::string core_of(const std::string& acc) {
    const size_t n = acc.size();
    if (n < 3 || acc[n - 2] != '/' || acc[n - 1] < '0' || acc[n - 1] > '9') return acc;
    for (size_t i = 0; i + 2 < n; i++) {
        const unsigned char c = (unsigned char)acc[i];
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') return acc;
    }
    return acc.substr(0, n - 2);
}

int main(int argc, char** argv) {
    perlc::LongOpts o(argc, argv, {"left", "right"}, {"sorted"});
    const std::string left = o.get("left"), right = o.get("right");
    const bool sorted = o.flag("sorted");
    if (left.empty() || left == "0" || right.empty() || right == "0") { fputs(USAGE, stderr); return 255; }
    fprintf(stderr, "-opening %s\n", left.c_str());
    FILE* lf = open_stats(left);
    fprintf(stderr, "-opening %s\n", right.c_str());
    FILE* rf = open_stats(right);
    fprintf(stderr, "-done opening files.\n");
    perlc::DelimReader lr(lf, fatal), rr(rf, fatal);
    const int la = lr.col("acc"), lm = lr.col("median_cov"), lu = lr.col("mean_cov"), ls = lr.col("stdev");
    const int ra = rr.col("acc"), rm = rr.col("median_cov"), ru = rr.col("mean_cov"), rs = rr.col("stdev");
    fputs("acc\tleft_acc\tleft_median_cov\tleft_mean_cov\tleft_stdev\tright_acc\tright_median_cov\tright_mean_cov\tright_stdev\t"
          "median_cov\tmean_cov\tstdev\n", stdout);
    std::vector<std::string> L, R;
    bool hl = lr.next(L), hr = rr.next(R);
    typedef perlc::DelimReader D;
    while (hl && hr) {
        const std::string& lacc = D::field(L, la);
        const std::string& racc = D::field(R, ra);
        const std::string core = core_of(lacc), rcore = core_of(racc);
        if (rcore != core) {
            if (!sorted)
                fatal("Error, core accs are not equivalent: [" + core + "] vs. [" + rcore + "] reads, and --sorted flag wasn't used here. at nbkc_merge_left_right_stats line 126.");
            if (lacc < racc) hl = lr.next(L); else hr = rr.next(R);
            continue;
        }
        const double med = perlc::perl_add(D::field(L, lm), D::field(R, rm)) / 2;
        const double mean = perlc::perl_add(D::field(L, lu), D::field(R, ru)) / 2;
        const double sd = perlc::perl_add(D::field(L, ls), D::field(R, rs)) / 2;
        std::string line = core;
        for (const std::string* s : {&lacc, &D::field(L, lm), &D::field(L, lu), &D::field(L, ls), &racc, &D::field(R, rm),
                                     &D::field(R, ru), &D::field(R, rs)}) { line.push_back('\t'); line += *s; }
        line.push_back('\t'); line += perlc::fmt_fixed(med, 1);
        line.push_back('\t'); line += perlc::fmt_fixed(mean, 1);
        line.push_back('\t'); line += perlc::fmt_fixed(sd, 1);
        line.push_back('\n');
        fwrite(line.data(), 1, line.size(), stdout);
        hl = lr.next(L);
        hr = rr.next(R);
    }
    if (fflush(stdout) != 0) { fprintf(stderr, "write failed\n"); return 1; }
    return 0;
}
