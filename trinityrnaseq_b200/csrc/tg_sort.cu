// Ascending sort of exported (packed k-mer, count) pairs: the dump is emitted in lexicographic k-mer order so
// that it is deterministic (jellyfish's own order is its internal hash order and nobody relies on it,
// SURVEY §8a J2).  Off the measured path; CUB radix sort from the CUDA toolkit.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "tg_internal.h"

namespace tg {

cudaError_t sort_pairs(uint64_t* d_keys, uint32_t* d_vals, uint64_t n, int k, cudaStream_t s) {
    if (n < 2) return cudaSuccess;
    if (n > 0x7FFFFFFFull) return cudaErrorInvalidValue;   // export is batched by the caller above this
    uint64_t* keys_alt = nullptr;
    uint32_t* vals_alt = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e;
    if ((e = cudaMallocAsync(&keys_alt, n * sizeof(uint64_t), s)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&vals_alt, n * sizeof(uint32_t), s)) != cudaSuccess) return e;
    cub::DoubleBuffer<uint64_t> kb(d_keys, keys_alt);
    cub::DoubleBuffer<uint32_t> vb(d_vals, vals_alt);
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 2 * k, s);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 1, s);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)n, 0, 2 * k, s);
    if (e == cudaSuccess && kb.Current() != d_keys) {
        e = cudaMemcpyAsync(d_keys, kb.Current(), n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_vals, vb.Current(), n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
    }
    if (tmp) cudaFreeAsync(tmp, s);
    cudaFreeAsync(keys_alt, s);
    cudaFreeAsync(vals_alt, s);
    return e;
}

}  // namespace tg

namespace tg {

// Locus order (tg_perread.cu): ascending sort of (32-bit read signature, read index).  `work` holds, in this order,
// sig[n] | idx[n] | sig_alt[n] | idx_alt[n] | CUB temporary storage; the caller has filled sig and idx.  Returns the
// device pointer of the sorted index array (one of the two idx buffers) in *d_sorted.
size_t locus_sort_bytes(uint64_t n) {
    size_t tmp = 0;
    cub::DoubleBuffer<uint32_t> kb(nullptr, nullptr), vb(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int)n, 0, 32, (cudaStream_t)0);
    return (size_t)n * 16 + ((tmp + 255) & ~(size_t)255) + 256;
}

cudaError_t locus_sort(void* work, size_t work_bytes, uint64_t n, const uint32_t** d_sorted, cudaStream_t s) {
    TimedLaunch timed("locus_sort", s);
    if (n > 0x7FFFFFF0ull) return cudaErrorInvalidValue;
    uint32_t* sig = (uint32_t*)work;
    uint32_t* idx = sig + n;
    uint32_t* sig_alt = idx + n;
    uint32_t* idx_alt = sig_alt + n;
    void* tmp = (void*)(((uintptr_t)(idx_alt + n) + 255) & ~(uintptr_t)255);
    size_t tmp_bytes = work_bytes - (size_t)((char*)tmp - (char*)work);
    cub::DoubleBuffer<uint32_t> kb(sig, sig_alt), vb(idx, idx_alt);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)n, 0, 32, s);
    *d_sorted = vb.Current();
    return e;
}

}  // namespace tg

namespace tg {

// Reads copied into locus order (tg_api.cu: tg_records_gather_locus_dev): lengths in the new order, exclusive scan
// (CUB), one warp per read copies its bytes (terminator included).
__global__ void __launch_bounds__(256)
k_gather_lengths(const uint64_t* __restrict__ offs, const uint32_t* __restrict__ order, uint64_t nreads, uint64_t* __restrict__ len) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < nreads) { const uint64_t r = order[i]; len[i] = offs[r + 1] - offs[r]; }
}
__global__ void __launch_bounds__(256)
k_gather_reads(const uint8_t* __restrict__ recs, const uint64_t* __restrict__ offs, const uint32_t* __restrict__ order,
               const uint64_t* __restrict__ out_off, uint64_t nreads, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= nreads) return;
    const uint64_t r = order[i], a = offs[r], n = offs[r + 1] - a, o = out_off[i];
    for (uint64_t j = lane; j < n; j += 32) out[o + j] = recs[a + j];
}

size_t gather_scratch_bytes(uint64_t nreads) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)nreads, (cudaStream_t)0);
    return (size_t)nreads * 16 + ((tmp + 255) & ~(size_t)255) + 256;
}

cudaError_t gather_reads(const uint8_t* d_recs, const uint64_t* d_offs, const uint32_t* d_order, uint64_t nreads, void* scratch,
                         size_t scratch_bytes, uint8_t* d_out, cudaStream_t s) {
    TimedLaunch timed("k_gather_reads", s);
    if (nreads == 0) return cudaSuccess;
    uint64_t* len = (uint64_t*)scratch;
    uint64_t* off = len + nreads;
    void* tmp = (void*)(((uintptr_t)(off + nreads) + 255) & ~(uintptr_t)255);
    size_t tmp_bytes = scratch_bytes - (size_t)((char*)tmp - (char*)scratch);
    k_gather_lengths<<<(unsigned)((nreads + 255) / 256), 256, 0, s>>>(d_offs, d_order, nreads, len);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, off, (int)nreads, s);
    if (e != cudaSuccess) return e;
    k_gather_reads<<<(unsigned)((nreads * 32 + 255) / 256), 256, 0, s>>>(d_recs, d_offs, d_order, off, nreads, d_out);
    return cudaGetLastError();
}

}  // namespace tg
