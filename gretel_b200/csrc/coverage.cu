// gretel-snpper on the GPU: per-position A,C,G,T counts over decoded alignments (gretel/snpper.py:30,
// pysam.count_coverage(quality_threshold=0, read_callback='nofilter') in the reference).
//
// The CPU side (bampack.cpp) decodes the BAM and ships, per wave, the aligned segments (one per M/=/X CIGAR
// operation: reference start, length, position of its first base in the 4-bit base array) and the 4-bit bases as
// they are stored in the BAM.  The kernel is a privatised histogram over a coordinate-sorted stream: a CTA takes a
// contiguous chunk of segments, which cover a narrow window of positions, counts them with shared-memory atomics
// into a [4][WIN] tile anchored at the chunk's first position and flushes the tile once; bases outside the tile
// (long reads, sparse coverage) go to HBM with integer reductions directly.
#include "hx_internal.cuh"

namespace {

constexpr int COV_WIN = 4096;            // positions per shared-memory tile (4 x 4096 x 4 B = 64 KB)
constexpr int COV_BLOCK = 512;

__global__ void __launch_bounds__(COV_BLOCK)
k_coverage(const int32_t *__restrict__ seg_start, const int64_t *__restrict__ seg_nib, const int32_t *__restrict__ seg_len,
           const uint8_t *__restrict__ seq4, int64_t n_seg, int64_t per_cta, int start0, int len,
           uint32_t *__restrict__ counts) {
    extern __shared__ uint32_t win[];                      // [4][COV_WIN]
    const int64_t c0 = (int64_t)blockIdx.x * per_cta;
    const int64_t c1 = c0 + per_cta < n_seg ? c0 + per_cta : n_seg;
    if (c0 >= c1) return;
    for (int i = threadIdx.x; i < 4 * COV_WIN; i += COV_BLOCK) win[i] = 0;
    const int base = seg_start[c0];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t s = c0 + warp; s < c1; s += COV_BLOCK / 32) {
        const int rs = seg_start[s], ln = seg_len[s];
        const int64_t q = seg_nib[s];
        for (int j = lane; j < ln; j += 32) {
            const int pos = rs + j;
            if (pos < start0 || pos >= start0 + len) continue;
            const int64_t n = q + j;
            const unsigned b = seq4[n >> 1];
            const unsigned nt = (n & 1) ? (b & 0xfu) : (b >> 4);         // "=ACMGRSVTWYHKDBN"
            const int code = nt == 1 ? 0 : nt == 2 ? 1 : nt == 4 ? 2 : nt == 8 ? 3 : -1;
            if (code < 0) continue;
            const int rel = pos - base;
            if (rel >= 0 && rel < COV_WIN) atomicAdd(&win[code * COV_WIN + rel], 1u);
            else atomicAdd(&counts[(size_t)code * len + (pos - start0)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * COV_WIN; i += COV_BLOCK) {
        const uint32_t v = win[i];
        if (!v) continue;
        const int code = i / COV_WIN, pos = base + (i - code * COV_WIN);
        if (pos >= start0 && pos < start0 + len) atomicAdd(&counts[(size_t)code * len + (pos - start0)], v);
    }
}

}  // namespace

// all pointers on the device; counts[4][len] accumulates
int hx_launch_coverage(const int32_t *d_seg_start, const int64_t *d_seg_nib, const int32_t *d_seg_len,
                       const uint8_t *d_seq4, int64_t n_seg, int32_t start0, int32_t len, uint32_t *d_counts,
                       cudaStream_t stream) {
    if (n_seg <= 0 || len <= 0) return HX_OK;
    static bool attr_set = false;
    if (!attr_set) {
        HX_CUDA(cudaFuncSetAttribute(k_coverage, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * COV_WIN * 4));
        attr_set = true;
    }
    // chunks of about 2048 segments: at deep coverage they span a few hundred positions
    const int64_t per_cta = 2048;
    const int64_t grid = (n_seg + per_cta - 1) / per_cta;
    k_coverage<<<(unsigned)grid, COV_BLOCK, 4 * COV_WIN * 4, stream>>>(d_seg_start, d_seg_nib, d_seg_len, d_seq4, n_seg,
                                                                      per_cta, start0, len, d_counts);
    HX_CUDA(cudaGetLastError());
    return HX_OK;
}
