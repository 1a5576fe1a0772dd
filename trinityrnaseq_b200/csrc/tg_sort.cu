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
