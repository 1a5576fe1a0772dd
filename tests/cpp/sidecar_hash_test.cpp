// CPU check of the host-side hand-off logic (trinityrnaseq_b200/host/tg_sidecar.hpp): the content hash that ties a
// .tgk sidecar to its FASTA must not depend on how the writer chunked the bytes, and must notice any single-byte edit.
#include <stdio.h>
#include <string>
#include <vector>
#include "tg_sidecar.hpp"

int main() {
    std::string text;
    unsigned long long x = 88172645463325252ull;
    for (int i = 0; i < 200000; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; text.push_back("ACGT>\n0123456789"[x % 16]); }
    tgside::TextHash whole;
    whole.update(text.data(), text.size());
    const unsigned long long ref = whole.digest();
    for (size_t step : {1u, 3u, 7u, 8u, 9u, 31u, 64u, 1000u, 65536u}) {
        tgside::TextHash h;
        for (size_t i = 0; i < text.size(); i += step) h.update(text.data() + i, std::min(step, text.size() - i));
        if (h.digest() != ref || h.total != text.size()) { printf("FAIL chunk %zu\n", step); return 1; }
    }
    for (size_t pos : {0ul, 1ul, 7ul, 8ul, 12345ul, text.size() - 1}) {
        std::string t2 = text;
        t2[pos] ^= 1;
        tgside::TextHash h;
        h.update(t2.data(), t2.size());
        if (h.digest() == ref) { printf("FAIL edit at %zu not noticed\n", pos); return 1; }
    }
    tgside::TextHash shorter, empty;
    shorter.update(text.data(), text.size() - 1);
    if (shorter.digest() == ref || empty.digest() == ref) { printf("FAIL length\n"); return 1; }
    if (sizeof(tgside::TgkHeader) != 40) { printf("FAIL header size %zu\n", sizeof(tgside::TgkHeader)); return 1; }
    printf("ok\n");
    return 0;
}
