// Device-wide scans as three small kernels (partials, spine, apply); header-only, used by the
// long-read ingestion (plane offsets, running max) and by the compact wire format (read offsets).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// ---- scans (three small kernels; T = value type, OP: 0 = sum, 1 = max) -------------------------
template <typename T, int OP>
__device__ __forceinline__ T scan_op(T a, T b) { return OP == 0 ? a + b : (a > b ? a : b); }

template <typename T, int OP, int ITEMS>
__global__ void __launch_bounds__(256)
k_scan_partials(const T *__restrict__ in, int64_t n, T *__restrict__ partials) {
    __shared__ T sh[8];
    const int64_t base = (int64_t)blockIdx.x * 256 * ITEMS;
    T acc = 0;
    for (int i = 0; i < ITEMS; ++i) {
        const int64_t j = base + (int64_t)i * 256 + threadIdx.x;
        if (j < n) acc = scan_op<T, OP>(acc, in[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = scan_op<T, OP>(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = 0;
        for (int w = 0; w < 8; ++w) t = scan_op<T, OP>(t, sh[w]);
        partials[blockIdx.x] = t;
    }
}

template <typename T, int OP>
__global__ void k_scan_spine(T *__restrict__ partials, int64_t nb, T *__restrict__ total) {
    if (threadIdx.x || blockIdx.x) return;               // nb is small (n / 4096)
    T run = 0;
    for (int64_t b = 0; b < nb; ++b) {
        const T v = partials[b];
        partials[b] = run;                               // exclusive
        run = scan_op<T, OP>(run, v);
    }
    if (total) *total = run;
}

// out[j] = exclusive scan (EXCL) or inclusive scan of in[0..j]
template <typename T, int OP, int ITEMS, bool EXCL>
__global__ void __launch_bounds__(256)
k_scan_apply(const T *in, int64_t n, const T *__restrict__ partials, T *out) {   // in may alias out
    __shared__ T sh[256];
    const int64_t base = (int64_t)blockIdx.x * 256 * ITEMS;
    T carry = partials[blockIdx.x];
    // each thread owns ITEMS consecutive elements
    const int64_t j0 = base + (int64_t)threadIdx.x * ITEMS;
    T loc[ITEMS];
    T sum = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        loc[i] = j0 + i < n ? in[j0 + i] : (T)0;
        sum = scan_op<T, OP>(sum, loc[i]);
    }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {                  // Hillis-Steele over the 256 thread sums
        T v = threadIdx.x >= o ? sh[threadIdx.x - o] : (T)0;
        __syncthreads();
        if (threadIdx.x >= o) sh[threadIdx.x] = scan_op<T, OP>(sh[threadIdx.x], v);
        __syncthreads();
    }
    T run = scan_op<T, OP>(carry, threadIdx.x ? sh[threadIdx.x - 1] : (T)0);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (j0 + i < n) {
            if (EXCL) out[j0 + i] = run;
            run = scan_op<T, OP>(run, loc[i]);
            if (!EXCL) out[j0 + i] = run;
        }
    }
}


}  // namespace
