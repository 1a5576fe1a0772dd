// Host-side text I/O of the three drop-in executables: FASTA readers that reproduce the observable quirks of the
// reference's three different readers, a record-buffer builder (sequence + '\n' terminator, the layout
// libtrinity_gpu consumes) and a buffered writer.  No k-mer arithmetic happens here.
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace tgio {

// ---- whole-file view (mmap for regular files, read() loop for pipes such as /dev/fd/0) -------------------
class FileView {
public:
    const char* data = nullptr;
    size_t size = 0;
    bool open(const std::string& path, std::string* err) {
        std::string p = path == "-" ? "/dev/fd/0" : path;
        fd_ = ::open(p.c_str(), O_RDONLY);
        if (fd_ < 0) { if (err) *err = "cannot open file " + path; return false; }
        struct stat st;
        if (fstat(fd_, &st) == 0 && S_ISREG(st.st_mode)) {
            size = (size_t)st.st_size;
            if (size == 0) { data = ""; return true; }
            void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) { if (err) *err = "mmap failed for " + path; return false; }
            madvise(m, size, MADV_SEQUENTIAL);
            data = (const char*)m; mapped_ = true;
            return true;
        }
        // stream: slurp
        size_t cap = 1 << 20;
        char* buf = (char*)malloc(cap);
        for (;;) {
            if (size == cap) { cap *= 2; buf = (char*)realloc(buf, cap); }
            ssize_t n = ::read(fd_, buf + size, cap - size);
            if (n < 0) { if (err) *err = "read failed for " + path; free(buf); return false; }
            if (n == 0) break;
            size += (size_t)n;
        }
        data = buf; owned_ = buf;
        return true;
    }
    ~FileView() {
        if (mapped_) munmap((void*)data, size);
        if (owned_) free(owned_);
        if (fd_ >= 0) ::close(fd_);
    }
private:
    int fd_ = -1;
    bool mapped_ = false;
    char* owned_ = nullptr;
};

// ---- record buffer -------------------------------------------------------------------------------------
// recs: each sequence followed by '\n'; offs[i] = start of record i, offs.back() = end of buffer.
// names: parallel blob of names (what the tool prints for the record).
struct RecordBatch {
    std::vector<char> recs;
    std::vector<uint64_t> offs{0};
    std::vector<char> names;
    std::vector<uint64_t> name_offs{0};
    size_t count() const { return offs.size() - 1; }
    void clear() { recs.clear(); offs.assign(1, 0); names.clear(); name_offs.assign(1, 0); }
    void end_record() { recs.push_back('\n'); offs.push_back(recs.size()); }
    void add_name(const char* p, size_t n) { names.insert(names.end(), p, p + n); name_offs.push_back(names.size()); }
    const char* seq(size_t i) const { return recs.data() + offs[i]; }
    size_t seq_len(size_t i) const { return (size_t)(offs[i + 1] - offs[i] - 1); }
    const char* name(size_t i) const { return names.data() + name_offs[i]; }
    size_t name_len(size_t i) const { return (size_t)(name_offs[i + 1] - name_offs[i]); }
};

inline const char* find_nl(const char* p, const char* end) {
    const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
    return q ? q : end;
}

// ---- Inchworm Fasta_reader + Fasta_entry (Inchworm/src/Fasta_reader.cpp:47-129, Fasta_entry.cpp:6-29) ------
// record = header line + every following line up to the next line that starts with '>'; ' ', '\t', '\n' removed
// from the sequence, upper-cased; text before the first header is ignored; a last line without '\n' is kept.
// next() appends the sequence to `out` (without terminator) and returns the raw header line (without '>').
class InchwormFastaReader {
public:
    InchwormFastaReader(const char* data, size_t size) : p_(data), end_(data + size) {
        // _init_reader: advance to the first line starting with '>'
        while (p_ < end_ && *p_ != '>') { const char* nl = find_nl(p_, end_); p_ = nl < end_ ? nl + 1 : end_; }
    }
    bool has_next() const { return p_ < end_; }
    // returns false at end of input
    bool next(const char** header, size_t* header_len, std::vector<char>& out) {
        if (p_ >= end_) return false;
        const char* nl = find_nl(p_, end_);
        *header = p_ + 1;                       // p_ points at '>'
        *header_len = (size_t)(nl - p_ - 1);
        p_ = nl < end_ ? nl + 1 : end_;
        while (p_ < end_ && *p_ != '>') {
            nl = find_nl(p_, end_);
            append_clean(p_, nl, out);
            p_ = nl < end_ ? nl + 1 : end_;
        }
        return true;
    }
    size_t consumed(const char* base) const { return (size_t)(p_ - base); }
private:
    static void append_clean(const char* a, const char* b, std::vector<char>& out) {
        const size_t o = out.size(), n = (size_t)(b - a);
        out.resize(o + n);
        char* w = out.data() + o;
        // branch-free upper-casing copy (vectorises); blanks are counted and squeezed out only if there were any
        size_t blanks = 0;
        for (size_t i = 0; i < n; i++) {
            const unsigned char c = (unsigned char)a[i];
            blanks += (size_t)((c == ' ') | (c == '\t'));
            w[i] = (char)(c - (((unsigned char)(c - 'a') < 26) << 5));
        }
        if (blanks) {
            char* d = w;
            for (size_t i = 0; i < n; i++) if (w[i] != ' ' && w[i] != '\t') *d++ = w[i];
            out.resize((size_t)(d - out.data()));
        }
    }
    const char* p_;
    const char* end_;
};

// accession = header up to the first space/tab after skipping leading ones (string_util::tokenize)
inline void accession_of(const char* h, size_t n, const char** acc, size_t* acc_len) {
    size_t i = 0;
    while (i < n && (h[i] == ' ' || h[i] == '\t')) i++;
    size_t j = i;
    while (j < n && h[j] != ' ' && h[j] != '\t') j++;
    *acc = h + i; *acc_len = j - i;
}

// ---- Chrysalis DNAStringStreamFast (Chrysalis/analysis/DNAVector.cc:1456-1501) -------------------------------
// name = the whole header line including '>'; the line after a header is ALWAYS sequence (even if it starts with
// '>'); further lines are appended verbatim until a line starting with '>'; a line that ends at EOF without '\n'
// makes the stream !good(): if it is the first sequence line the record is dropped, otherwise just that line is lost.
class DnaStreamReader {
public:
    DnaStreamReader(const char* data, size_t size) : p_(data), end_(data + size) {
        good_ = getline();                                    // ReadStream: first getline
        while (good_ && !(ll_ > 0 && *lp_ == '>')) good_ = getline();
    }
    // appends sequence (no terminator) to out; name points into the file
    bool next(const char** name, size_t* name_len, std::vector<char>& out) {
        if (!good_) return false;
        *name = lp_; *name_len = ll_;
        good_ = getline();
        if (!good_) return false;                             // namev.pop_back(): record dropped
        out.insert(out.end(), lp_, lp_ + ll_);
        good_ = getline();
        while (good_ && !(ll_ > 0 && *lp_ == '>')) {
            out.insert(out.end(), lp_, lp_ + ll_);
            good_ = getline();
        }
        return true;
    }
private:
    // std::getline semantics: returns stream.good() afterwards (false once a read touched EOF)
    bool getline() {
        if (p_ >= end_) { lp_ = end_; ll_ = 0; return false; }
        const char* nl = (const char*)memchr(p_, '\n', (size_t)(end_ - p_));
        if (!nl) { lp_ = p_; ll_ = (size_t)(end_ - p_); p_ = end_; return false; }
        lp_ = p_; ll_ = (size_t)(nl - p_); p_ = nl + 1;
        return true;
    }
    const char* p_;
    const char* end_;
    const char* lp_ = nullptr;
    size_t ll_ = 0;
    bool good_ = false;
};

// DNAStringStreamFast::formatReadNameString (DNAVector.cc:1504-1514)
inline void format_read_name(const char* n, size_t len, std::string& out) {
    out.assign(n, len);
    while (!out.empty() && out[0] == ' ') out.erase(0, 1);
    for (auto& c : out) if (c == ' ') c = '_';
    while (!out.empty() && out.back() == ' ') out.pop_back();
}

// ---- Chrysalis vecDNAVector::Read(f,false,false,true,..) (DNAVector.cc:856-971; FlatFileParser/Tokenize) ---
// lines are split on ' ' and '\t'; a line whose first token starts with '>' opens a record named by its tokens
// joined with '_'; any other non-empty line contributes its FIRST token to the sequence; upper-cased; the last
// line is lost when it lacks '\n' (ParseLine returns false once EOF was touched).
inline void read_bundles(const char* data, size_t size, RecordBatch& rb, std::vector<std::string>& names) {
    rb.clear(); names.clear();
    const char* p = data; const char* end = data + size;
    bool open = false;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!nl) break;                                         // unterminated last line: dropped
        // tokenise
        const char* q = p;
        bool first = true; bool is_header = false;
        std::string name;
        while (q < nl) {
            while (q < nl && (*q == ' ' || *q == '\t')) q++;
            if (q >= nl) break;
            const char* t = q;
            while (q < nl && *q != ' ' && *q != '\t') q++;
            if (first) {
                is_header = *t == '>';
                if (is_header) name.assign(t, (size_t)(q - t));
                else if (open) {
                    size_t o = rb.recs.size();
                    rb.recs.insert(rb.recs.end(), t, q);
                    for (size_t i = o; i < rb.recs.size(); i++)
                        if (rb.recs[i] >= 'a' && rb.recs[i] <= 'z') rb.recs[i] = (char)(rb.recs[i] - 32);
                }
                first = false;
                if (!is_header) break;
            } else {
                name += '_'; name.append(t, (size_t)(q - t));
            }
        }
        if (is_header) {
            if (open) rb.end_record();
            names.push_back(name);
            open = true;
        }
        p = nl + 1;
    }
    if (open) rb.end_record();
}

// ---- buffered output -------------------------------------------------------------------------------------
class OutBuf {
public:
    explicit OutBuf(int fd, size_t cap = 8u << 20) : fd_(fd) { buf_.reserve(cap); cap_ = cap; }
    ~OutBuf() { flush(); }
    void put(const char* p, size_t n) {
        if (buf_.size() + n > cap_) flush();
        if (n > cap_) { write_all(p, n); return; }
        buf_.insert(buf_.end(), p, p + n);
    }
    void put(const std::string& s) { put(s.data(), s.size()); }
    void putc(char c) { if (buf_.size() + 1 > cap_) flush(); buf_.push_back(c); }
    void put_uint(unsigned long long v) {
        char t[24]; int n = 0;
        do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        char r[24]; for (int i = 0; i < n; i++) r[i] = t[n - 1 - i];
        put(r, (size_t)n);
    }
    void put_int(long long v) { if (v < 0) { putc('-'); put_uint((unsigned long long)(-v)); } else put_uint((unsigned long long)v); }
    bool flush() {
        bool ok = write_all(buf_.data(), buf_.size());
        buf_.clear();
        return ok;
    }
    bool failed() const { return failed_; }
private:
    bool write_all(const char* p, size_t n) {
        while (n) {
            ssize_t w = ::write(fd_, p, n);
            if (w < 0) { failed_ = true; return false; }
            p += w; n -= (size_t)w;
        }
        return true;
    }
    int fd_;
    size_t cap_;
    std::vector<char> buf_;
    bool failed_ = false;
};

}  // namespace tgio
