// CPU check of trinityrnaseq_b200/csrc/tg_minimizer.cuh: the fast path (all windows of a read from sliding minima over
// strips, read orientation, canonical keys of either strand) must give exactly the home (h, j) that the slow path
// computes from the key alone -- for every k, strip width, tie pattern (tandem repeats, homopolymers, hairpins).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../trinityrnaseq_b200/csrc/tg_minimizer.cuh"

using namespace tg;

static unsigned code_of(char c) { return ((unsigned)c >> 1 ^ (unsigned)c >> 2) & 3u; }

template <int PER>
static long check_read(const std::string& s, int k, bool canonical) {
    const int L = (int)s.size(), m = mm_len(k);
    if (L < k) return 0;
    const int nwin = L - k + 1, nmm = L - m + 1;
    // planes of the whole read in 32-base chunks
    std::vector<unsigned> P0((L + 31) / 32 + 2, 0), P1((L + 31) / 32 + 2, 0);
    for (int i = 0; i < L; i++) {
        const unsigned c = code_of(s[i]);
        P0[i >> 5] |= (c & 1u) << (i & 31);
        P1[i >> 5] |= (c >> 1) << (i & 31);
    }
    std::vector<unsigned> hx(nmm + 64, 0xDEADBEEFu);      // garbage past the end must never matter
    const unsigned mm = bits_mask(m), mk = bits_mask(k);
    for (int q = 0; q < nmm; q++)
        hx[q] = mmer_hash(funnel_r(P0[q >> 5], P0[(q >> 5) + 1], q & 31) & mm, funnel_r(P1[q >> 5], P1[(q >> 5) + 1], q & 31) & mm, m);
    long checked = 0;
    for (int s0 = 0; s0 < nwin; s0 += PER) {
        unsigned strip[PER + HOME_SLOTS - 1], vl[PER], vr[PER];
        for (int i = 0; i < PER + HOME_SLOTS - 1; i++) strip[i] = hx[s0 + i];
        strip_minimizers<PER>(strip, vl, vr);
        for (int i = 0; i < PER && s0 + i < nwin; i++) {
            const int p = s0 + i;
            const unsigned f0 = funnel_r(P0[p >> 5], P0[(p >> 5) + 1], p & 31) & mk;
            const unsigned f1 = funnel_r(P1[p >> 5], P1[(p >> 5) + 1], p & 31) & mk;
            const unsigned r0 = rc_plane_n(f0, k), r1 = rc_plane_n(f1, k);
            const unsigned long long kf = ((unsigned long long)f1 << 32) | f0, kr = ((unsigned long long)r1 << 32) | r0;
            const bool is_rc = canonical && kr < kf;
            unsigned j_fast, h_slow, j_slow;
            const unsigned sp = strip_pick(vl[i], vr[i], i, is_rc, j_fast);
            const unsigned h_fast = hx[s0 + sp];
            key_home(is_rc ? r0 : f0, is_rc ? r1 : f1, k, h_slow, j_slow);
            if (h_fast != h_slow || j_fast != j_slow) {
                fprintf(stderr, "MISMATCH k=%d PER=%d canonical=%d window %d of %s: fast (%08x,%u) slow (%08x,%u)\n", k, PER,
                        (int)canonical, p, s.c_str(), h_fast, j_fast, h_slow, j_slow);
                exit(1);
            }
            // the reverse complement of the key must have the mirrored home
            unsigned h2, j2;
            key_home(is_rc ? f0 : r0, is_rc ? f1 : r1, k, h2, j2);
            (void)h2; (void)j2;       // (not required to agree under hash ties between different m-mers)
            checked++;
        }
    }
    return checked;
}

int main() {
    srand(12345);
    long n = 0;
    std::vector<std::string> reads;
    for (int r = 0; r < 300; r++) {
        const int L = 20 + rand() % 200;
        std::string s(L, 'A');
        for (auto& c : s) c = "ACGT"[rand() & 3];
        reads.push_back(s);
    }
    // tie-heavy reads: homopolymers, short-period tandem repeats, hairpins (a sequence followed by its reverse complement)
    reads.push_back(std::string(120, 'A'));
    reads.push_back(std::string(90, 'T'));
    for (int period = 1; period <= 9; period++) {
        std::string unit(period, 'A');
        for (auto& c : unit) c = "ACGT"[rand() & 3];
        std::string s;
        while (s.size() < 150) s += unit;
        reads.push_back(s);
        // ... with a random prefix and suffix so that repeats enter and leave the windows
        std::string t = reads[period] + s + reads[period + 20];
        reads.push_back(t);
    }
    for (int r = 0; r < 40; r++) {
        std::string half = reads[r].substr(0, 10 + rand() % 40), rc(half.rbegin(), half.rend());
        for (auto& c : rc) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
        reads.push_back(half + rc + half);
    }
    const int ks[] = {8, 9, 15, 20, 21, 24, 25, 31, 32};
    for (int k : ks)
        for (const auto& s : reads)
            for (int canonical = 0; canonical < 2; canonical++) {
                n += check_read<1>(s, k, canonical); n += check_read<2>(s, k, canonical); n += check_read<3>(s, k, canonical);
                n += check_read<4>(s, k, canonical); n += check_read<5>(s, k, canonical); n += check_read<6>(s, k, canonical);
                n += check_read<7>(s, k, canonical); n += check_read<8>(s, k, canonical);
            }
    // k < 8: only the slow path exists; it must be strand-consistent in its hash
    for (int k = 1; k < 8; k++) {
        unsigned h, j;
        key_home(0x15u & bits_mask(k), 0x0Au & bits_mask(k), k, h, j);
        if ((int)j >= mm_win(k)) { fprintf(stderr, "slot out of range for k=%d\n", k); return 1; }
    }
    printf("minimizer fast == slow on %ld windows\n", n);
    return 0;
}
