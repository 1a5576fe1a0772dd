// Synthetic RNA-seq style paired reads generated ON THE DEVICE (bench / test utility, not a product path).
// Pure integer arithmetic (splitmix64 counters), so a given (transcriptome, seed) always yields the same bytes.
// Shape follows SURVEY §8(d): transcripts picked by expression weight, fragment ~ N(mean, sd) (Irwin-Hall
// of four uniforms), left mate = fragment prefix, right mate = reverse complement of the fragment suffix,
// optional strand coin-flip, i.i.d. substitutions and N calls.  Output = record buffer: all left mates then all
// right mates, each `read_len` bases + '\n' (fixed stride read_len + 1).
#include "tg_internal.h"

namespace tg {

__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint8_t comp_base(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

__global__ void __launch_bounds__(256)
k_synth_reads(const uint8_t* __restrict__ tx, const uint64_t* __restrict__ tx_offs, const uint64_t* __restrict__ tx_cum,
              uint32_t ntx, uint64_t npairs, int read_len, int frag_mean, int frag_sd, uint32_t err_per_million,
              uint32_t n_per_million, uint64_t seed, int stranded, uint8_t* __restrict__ recs) {
    const int lane = threadIdx.x & 31;
    const uint64_t rec = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (rec >= 2 * npairs) return;
    const int mate = rec >= npairs;
    const uint64_t pair = mate ? rec - npairs : rec;
    // per-pair draws (identical in every lane and for both mates)
    const unsigned long long s0 = splitmix(seed ^ (pair * 0xD1342543DE82EF95ull));
    const unsigned long long u_tx = splitmix(s0 + 1), u_len = splitmix(s0 + 2), u_pos = splitmix(s0 + 3),
                             u_str = splitmix(s0 + 4);
    uint32_t lo = 0, hi = ntx;   // first transcript whose cumulative threshold exceeds u_tx
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (tx_cum[mid] > u_tx) hi = mid; else lo = mid + 1; }
    if (lo >= ntx) lo = ntx - 1;
    const uint64_t t0 = tx_offs[lo];
    const int tlen = (int)(tx_offs[lo + 1] - t0);
    const long long ih = (long long)((u_len & 0xFFFF) + ((u_len >> 16) & 0xFFFF) + ((u_len >> 32) & 0xFFFF) + (u_len >> 48)) - 131072;
    int flen = frag_mean + (int)(ih * frag_sd / 37837);
    if (flen < read_len) flen = read_len;
    if (flen > tlen) flen = tlen;
    int rl = read_len < flen ? read_len : flen;
    const int start = (int)(u_pos % (unsigned long long)(tlen - flen + 1));
    const bool flip = !stranded && (u_str & 1ull);
    // mate 0 reads the fragment forward from its start, mate 1 reads its reverse complement from its end;
    // a flipped fragment swaps the roles
    const bool rev = (mate == 1) != flip;
    uint8_t* out = recs + rec * (uint64_t)(read_len + 1);
    for (int i = lane; i <= read_len; i += 32) {
        uint8_t c;
        if (i == read_len) c = '\n';
        else if (i >= rl) c = 'N';
        else {
            c = rev ? comp_base(tx[t0 + start + flen - 1 - i]) : tx[t0 + start + i];
            const unsigned long long e = splitmix(s0 ^ ((unsigned long long)(mate * 1000003 + i + 17) * 0xA24BAED4963EE407ull));
            const uint32_t roll = (uint32_t)(e % 1000000ull);
            if (roll < err_per_million) {
                const uint32_t sub = (uint32_t)((e >> 40) % 3ull);
                const char alt[4] = {'A', 'C', 'G', 'T'};
                uint32_t code = base_code(c);
                c = alt[(code + 1 + sub) & 3u];
            } else if (roll < err_per_million + n_per_million) c = 'N';
        }
        out[i] = c;
    }
}

cudaError_t launch_synth_reads(const uint8_t* d_tx, const uint64_t* d_tx_offs, const uint64_t* d_tx_cum, uint32_t ntx,
                               uint64_t npairs, int read_len, int frag_mean, int frag_sd, uint32_t err_per_million,
                               uint32_t n_per_million, uint64_t seed, int stranded, uint8_t* d_recs, cudaStream_t s) {
    if (npairs == 0) return cudaSuccess;
    const uint64_t warps = 2 * npairs;
    const uint64_t blocks = (warps * 32 + 255) / 256;
    k_synth_reads<<<(unsigned)blocks, 256, 0, s>>>(d_tx, d_tx_offs, d_tx_cum, ntx, npairs, read_len, frag_mean, frag_sd,
                                                   err_per_million, n_per_million, seed, stranded, d_recs);
    return cudaGetLastError();
}

}  // namespace tg
