// host/fmt_float.hpp against printf("%g") (what the reference's ostream prints): every kind of float, byte for byte
#include <stdint.h>
#include <string.h>

#include <random>

#include "fmt_float.hpp"

int main() {
    std::mt19937_64 rng(20251017);
    size_t bad = 0, n = 0;
    char a[64], b[64];
    auto check = [&](float f) {
        int n1;
        if (isnan(f)) n1 = sprintf(a, signbit(f) ? "-nan" : "nan"); else n1 = sprintf(a, "%g", (double)f);
        const int n2 = fmt_float(b, f);
        n++;
        if (n1 != n2 || memcmp(a, b, (size_t)n1) != 0) { if (bad++ < 10) { b[n2] = 0; fprintf(stderr, "DIFF %a: '%s' vs '%s'\n", (double)f, a, b); } }
    };
    const float special[] = {0.0f, -0.0f, 1.0f, 0.5f, 1e-5f, 9.99999e-5f, 1e-4f, 999999.0f, 999999.5f, 1e6f, 1.5e6f, 123456.5f, 1234565.0f,
                             3.4028235e38f, 1.17549435e-38f, 1e-45f, INFINITY, -INFINITY, 1.33378e9f, 76.7308f, 32.3844f, 0.464829f};
    for (float f : special) { check(f); check(-f); }
    uint32_t nanbits = 0xFFC00000u; float fn; memcpy(&fn, &nanbits, 4); check(fn);
    nanbits = 0x7FC00000u; memcpy(&fn, &nanbits, 4); check(fn);
    for (int i = 0; i < 3000000; i++) {
        const uint32_t u = (uint32_t)rng();
        float f;
        if (i % 3 == 0) memcpy(&f, &u, 4);                                                  // any bit pattern
        else if (i % 3 == 1) f = (float)(u % 100000) / (float)((u >> 20) % 977 + 1);         // means
        else f = sqrtf((float)(u % 1000000) / (float)((u >> 20) % 76 + 1));                 // standard deviations
        check(f);
    }
    printf("%zu floats, %zu differences\n", n, bad);
    return bad ? 1 : 0;
}
