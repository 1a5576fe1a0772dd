// The few pieces of Perl semantics the normalisation helper scripts depend on (util/support_scripts/nbkc_*.pl,
// PerlLib/DelimParser.pm), restated so that their C++ drop-ins print the same bytes:
//   * string -> number the way Perl numifies a scalar (leading blanks, sign, decimal / exponent, inf / nan spellings,
//     trailing garbage ignored, no hex);
//   * rand() after srand(seed): Perl's own drand48 (perl util.c, Perl_drand48_r: 48-bit LCG a = 0x5DEECE66D, c = 0xB,
//     state = seed << 16 | 0x330E, result = state / 2^48) -- identical on every platform since perl 5.20;
//   * sprintf("%.Nf") incl. Perl's spelling of non-finite values ("NaN", "Inf", "-Inf");
//   * DelimParser::Reader: header line names the columns, rows are split on the delimiter with Perl's split()
//     semantics, a row whose field count differs from the header's is fatal, a row that is the single character "0"
//     without newline ends the input (Perl truthiness of the line).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include <string>
#include <vector>

namespace perlc {

inline double numify(const std::string& s) {
    const char* p = s.c_str();
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v') p++;
    const char* q = p;
    if (*q == '+' || *q == '-') q++;
    if (q[0] == '0' && (q[1] == 'x' || q[1] == 'X')) return (*p == '-') ? -0.0 : 0.0;      // Perl does not read hex strings
    char* end = nullptr;
    const double v = strtod(p, &end);
    if (end == p) return 0.0;
    return v;
}

// $a + $b on two strings, the way pp_add decides between integer and floating addition (PERL_PRESERVE_IVUV).  The
// results differ only in the SIGN OF A ZERO SUM, which "%.1f" prints: the right operand is coerced first; if it is
// integer-like and so is the left one the sum is an integer (so never -0); a plain integer literal that was coerced to
// an integer first has lost its sign ("-0" is 0), one that is only ever read as a float keeps it ("-0" is -0.0); a
// decimal-point spelling ("-0.0") is never integer-like.  Verified against perl 5.38 for every pair of zero spellings
// (tests/test_nbkc_tools.py).
struct PerlNum {
    double nv = 0;          // value read as a float (atof)
    bool int_literal = false, ivable = false;
    long long iv = 0;
};
inline PerlNum classify(const std::string& s) {
    PerlNum r;
    r.nv = numify(s);
    const char* p = s.c_str();
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v') p++;
    const char* q = p;
    if (*q == '+' || *q == '-') q++;
    const char* d = q;
    while (*d >= '0' && *d <= '9') d++;
    const bool has_digits = d > q;
    const bool frac = *d == '.';
    const char* e = d;
    if (frac) { e++; while (*e >= '0' && *e <= '9') e++; }
    const bool expo = has_digits && (*e == 'e' || *e == 'E');
    if (has_digits && !frac && !expo && (d - q) <= 18) {               // plain integer literal (trailing garbage ignored)
        r.int_literal = r.ivable = true;
        r.iv = strtoll(p, nullptr, 10);
        r.nv = (double)r.iv;                                           // as seen AFTER integer coercion; see perl_add
    } else if (has_digits && frac && !expo && (e - d) > 1) {
        r.ivable = false;                                              // "12.50", "-0.0": private integer flag only
    } else if (isfinite(r.nv) && fabs(r.nv) < 9.2e18 && (double)(long long)r.nv == r.nv) {
        r.ivable = true;                                               // "1e3", "-0e0", "7." ...: integral float
        r.iv = (long long)r.nv;
    }
    return r;
}
inline bool plain_positive(const std::string& s) {       // digits[.digits][e[+-]digits], first char a digit 1-9 or "0." ...
    if (s.empty() || s[0] < '0' || s[0] > '9') return false;
    bool nonzero = false;
    for (char c : s) {
        if (c >= '1' && c <= '9') nonzero = true;
        else if (!(c == '0' || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-')) return false;
    }
    return nonzero;
}
inline double perl_add(const std::string& a, const std::string& b) {
    // ordinary positive numbers: integer and floating addition agree (below 2^53) and no zero can result
    if (a.size() <= 15 && b.size() <= 15 && plain_positive(a) && plain_positive(b)) return strtod(a.c_str(), nullptr) + strtod(b.c_str(), nullptr);
    const PerlNum A = classify(a), B = classify(b);
    if (B.ivable && A.ivable) {
        long long sum;
        if (!__builtin_add_overflow(A.iv, B.iv, &sum)) return (double)sum;       // integer sum: +0 when zero
        return (double)A.iv + (double)B.iv;
    }
    if (B.ivable) return numify(a) + (B.int_literal ? (double)B.iv : B.nv);       // right coerced first: "-0" became 0
    return numify(a) + numify(b);                                                 // both only ever read as floats
}

struct Drand48 {
    uint64_t x;
    explicit Drand48(uint32_t seed) : x(((uint64_t)seed << 16) + 0x330Eull) {}
    double next() {
        x = (x * 0x5DEECE66Dull + 0xBull) & 0xFFFFFFFFFFFFull;
        return ldexp((double)x, -48);
    }
};

inline std::string fmt_fixed(double v, int prec) {
    if (isnan(v)) return "NaN";
    if (isinf(v)) return v < 0 ? "-Inf" : "Inf";
    char buf[512];
    snprintf(buf, sizeof buf, "%.*f", prec, v);
    return buf;
}

// Perl split(/\t/, $line): trailing empty fields are dropped
inline void split_tab(const std::string& line, std::vector<std::string>& out) {
    out.clear();
    size_t a = 0;
    while (true) {
        size_t b = line.find('\t', a);
        if (b == std::string::npos) { out.push_back(line.substr(a)); break; }
        out.push_back(line.substr(a, b - a));
        a = b + 1;
    }
    while (!out.empty() && out.back().empty()) out.pop_back();
}

class DelimReader {
public:
    std::vector<std::string> columns;
    // fatal(msg) must not return
    DelimReader(FILE* f, void (*fatal)(const std::string&)) : f_(f), fatal_(fatal) {
        std::string h;
        if (!getline(h)) fatal_("Error, no header row read.");
        if (!h.empty() && h.back() == '\n') h.pop_back();
        if (h.empty() || h == "0") fatal_("Error, no header row read.");
        split_tab(h, columns);
    }
    int col(const char* name) const {
        for (size_t i = 0; i < columns.size(); i++) if (columns[i] == name) return (int)i;
        return -1;
    }
    // false at end of input
    bool next(std::vector<std::string>& fields) {
        std::string line;
        if (!getline(line)) return false;
        if (line == "0") return false;                     // `unless ($line)`: the string "0" is false in Perl
        split_tab(line, fields);
        if (!fields.empty() && !fields.back().empty() && fields.back().back() == '\n') fields.back().pop_back();
        if (fields.size() != columns.size())
            fatal_("Error, line: [" + line + "] is lacking " + std::to_string(columns.size()) + " fields");
        return true;
    }
    static const std::string& field(const std::vector<std::string>& f, int c) {
        static const std::string empty;
        return c >= 0 && (size_t)c < f.size() ? f[(size_t)c] : empty;
    }
private:
    bool getline(std::string& out) {
        const ssize_t n = ::getline(&buf_, &cap_, f_);          // POSIX getline: one pass, reused buffer
        if (n <= 0) { out.clear(); return false; }
        out.assign(buf_, (size_t)n);
        return true;
    }
    char* buf_ = nullptr;
    size_t cap_ = 0;
    FILE* f_;
    void (*fatal_)(const std::string&);
};

// Getopt::Long as the two scripts use it: --name value | --name=value; `=i` values must be integers
struct LongOpts {
    std::vector<std::pair<std::string, std::string>> kv;
    std::vector<std::string> flags;
    LongOpts(int argc, char** argv, const std::vector<std::string>& with_value, const std::vector<std::string>& bare) {
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            if (a.size() < 3 || a[0] != '-' || a[1] != '-') continue;            // pass_through: everything else is ignored
            std::string name = a.substr(2), val;
            bool has_val = false;
            size_t eq = name.find('=');
            if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
            bool takes = false, is_flag = false;
            for (auto& w : with_value) if (w == name) takes = true;
            for (auto& b : bare) if (b == name) is_flag = true;
            if (takes) {
                if (!has_val) { if (i + 1 >= argc) continue; val = argv[++i]; }
                kv.push_back({name, val});
            } else if (is_flag) {
                flags.push_back(name);
            }
        }
    }
    bool has(const std::string& n) const { for (auto& p : kv) if (p.first == n) return true; return false; }
    std::string get(const std::string& n) const { std::string v; for (auto& p : kv) if (p.first == n) v = p.second; return v; }
    bool flag(const std::string& n) const { for (auto& f : flags) if (f == n) return true; return false; }
    static bool is_int(const std::string& s) {
        size_t i = (!s.empty() && (s[0] == '-' || s[0] == '+')) ? 1 : 0;
        if (i >= s.size()) return false;
        for (; i < s.size(); i++) if (s[i] < '0' || s[i] > '9') return false;
        return true;
    }
};

}  // namespace perlc
