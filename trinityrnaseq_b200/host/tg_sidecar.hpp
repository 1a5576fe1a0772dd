// Binary hand-off between `jellyfish dump` and the tools that read its FASTA.
//
// The reference moves the k-mer table between processes as text: `jellyfish dump -L n db > kmers.fa`
// (Trinity:2625, util/insilico_read_normalization.pl:641), re-parsed record by record into a hash by
// fastaToKmerCoverageStats --kmers (Inchworm/src/fastaToKmerCoverageStats.cpp:181-228) and inchworm --kmers
// (Inchworm/src/IRKE.cpp:81-154).  The FASTA stays (unchanged inchworm still reads it); next to it `dump` leaves
// `<file>.tgk`: the same records as packed (key, count) pairs.  A consumer uses the sidecar only if it provably
// describes the FASTA it was asked to read -- same byte length and same 64-bit content hash -- so editing, replacing or
// truncating the FASTA silently falls back to parsing the text.  TRINITY_GPU_NO_SIDECAR=1 disables both ends.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>

namespace tgside {

const char TGK_MAGIC[8] = {'T', 'G', 'K', 'M', 'E', 'R', '1', '\n'};

struct TgkHeader {
    char magic[8];
    uint32_t k;
    uint32_t reserved;
    uint64_t n;            // records: n u64 packed k-mers (A=0 C=1 G=2 T=3, first base most significant), then n u32 counts
    uint64_t text_bytes;   // length of the FASTA this describes
    uint64_t text_hash;    // TextHash of its bytes
};

// streaming 64-bit hash, independent of how the bytes are chunked
struct TextHash {
    uint64_t h = 0x9E3779B97F4A7C15ull, word = 0, total = 0;
    unsigned fill = 0;
    inline void mix(uint64_t w) { h = (h ^ w) * 0xff51afd7ed558ccdULL; h ^= h >> 32; }
    void update(const char* p, size_t n) {
        total += n;
        while (n && fill) { word |= (uint64_t)(unsigned char)*p++ << (8 * fill); n--; if (++fill == 8) { mix(word); word = 0; fill = 0; } }
        for (; n >= 8; n -= 8, p += 8) { uint64_t w; memcpy(&w, p, 8); mix(w); }
        for (; n; n--) { word |= (uint64_t)(unsigned char)*p++ << (8 * fill); if (++fill == 8) { mix(word); word = 0; fill = 0; } }
    }
    uint64_t digest() const { uint64_t x = h; if (fill) x = (x ^ word) * 0xc4ceb9fe1a85ec53ULL; x ^= total; x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; return x; }
};

inline bool disabled() { const char* e = getenv("TRINITY_GPU_NO_SIDECAR"); return e && *e && *e != '0'; }

// path of the regular file behind fd, empty if fd is a pipe / terminal / unknown, or not positioned at its start
inline std::string regular_file_behind(int fd) {
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return "";
    if (lseek(fd, 0, SEEK_CUR) != 0) return "";                 // appending (>>) : the file holds more than our dump
    char link[64], path[4096];
    snprintf(link, sizeof link, "/proc/self/fd/%d", fd);
    ssize_t n = readlink(link, path, sizeof path - 1);
    if (n <= 0) return "";
    path[n] = 0;
    if (path[0] != '/' || strstr(path, " (deleted)")) return "";
    return path;
}

}  // namespace tgside
