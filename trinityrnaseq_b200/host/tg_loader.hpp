// Shared glue of the three executables: error handling around the C ABI, device selection, packed k-mer <-> text.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "trinity_gpu.h"

namespace tgh {

// Non-zero exit + message on stderr is the whole error convention of the CLI boundary
// (PerlLib/Pipeliner.pm:176-187).  There is no CPU fallback: without a GPU the tools fail here.
[[noreturn]] inline void die(int code, const char* what) {
    fprintf(stderr, "ERROR: %s: %s\n", what, tg_last_error());
    exit(code);
}
#define TGC(call) do { if ((call) != TG_OK) tgh::die(3, #call); } while (0)

// device selection comes from the environment only (TRINITY_GPU=<index>), never from new argv
inline tg_ctx* open_device() {
    int dev = 0;
    if (const char* e = getenv("TRINITY_GPU")) dev = atoi(e);
    tg_ctx* ctx = nullptr;
    if (tg_init(dev, &ctx) != TG_OK) die(3, "tg_init");
    return ctx;
}

inline int base_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}
// returns false when the string holds a non-ACGT character
inline bool pack_kmer(const char* s, int k, uint64_t* out) {
    uint64_t v = 0;
    for (int i = 0; i < k; i++) {
        int c = base_code(s[i]);
        if (c < 0) return false;
        v = (v << 2) | (uint64_t)c;
    }
    *out = v;
    return true;
}
inline void unpack_kmer(uint64_t v, int k, char* out) {
    for (int i = k - 1; i >= 0; i--) { out[i] = "ACGT"[v & 3]; v >>= 2; }
}

}  // namespace tgh
