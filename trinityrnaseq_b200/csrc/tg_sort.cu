// Ascending sort of exported (packed k-mer, count) pairs: the dump is emitted in lexicographic k-mer order so
// that it is deterministic (jellyfish's own order is its internal hash order and nobody relies on it,
// SURVEY §8a J2).  Off the measured path; CUB radix sort from the CUDA toolkit.
#include <cub/device/device_radix_sort.cuh>
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
