// K1: pair-expansion ingestion of packed reads into the banded Hansel counts.
// Replaces gretel/util.py:226-286 (+ Hansel.add_observation) of the reference.
//
// Two kernels:
//   k1_pairs_red    generic: one warp per read, lanes over the linearised (i,j)
//                   triangle, one fire-and-forget integer reduction (RED) per pair
//                   into the L2-resident band.  Any read order, any k.
//   k1_bitsliced    (see below) rank-sorted short reads: 32 reads per warp are
//                   transposed into per-site allele bit-planes with warp ballots; a
//                   pair of sites then costs AND+POPC per (a,b) instead of one atomic
//                   per read, accumulated in registers and flushed once per tile.
#include "hx_internal.cuh"

namespace {

__device__ __forceinline__ unsigned long long warp_sum_ull(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level accumulation of the four ingestion totals into global memory.
template <int BLOCK>
__device__ __forceinline__ void flush_totals(unsigned long long t0, unsigned long long t1,
                                             unsigned long long t2, unsigned long long t3,
                                             unsigned long long *totals) {
    __shared__ unsigned long long sh[4];
    if (threadIdx.x < 4) sh[threadIdx.x] = 0;
    __syncthreads();
    t0 = warp_sum_ull(t0); t1 = warp_sum_ull(t1); t2 = warp_sum_ull(t2); t3 = warp_sum_ull(t3);
    if ((threadIdx.x & 31) == 0) {
        if (t0) atomicAdd(&sh[0], t0);
        if (t1) atomicAdd(&sh[1], t1);
        if (t2) atomicAdd(&sh[2], t2);
        if (t3) atomicAdd(&sh[3], t3);
    }
    __syncthreads();
    if (threadIdx.x < 4 && sh[threadIdx.x]) atomicAdd(&totals[threadIdx.x], sh[threadIdx.x]);
}

// The per-pair rules of util.py:254-281 for one (i,j) of one read.
__device__ __forceinline__ void add_pair(uint32_t *__restrict__ cnt, int N, int64_t W, int rk, int i,
                                         int j, unsigned a, unsigned b, unsigned long long &sent) {
    const int pi = rk + i + 1, pj = rk + j + 1;
    atomicAdd(cnt + hx_cell_off(W, pi, pj) + a * HX_NSYM + b, 1u);            // :267,274,280
    if (i == 0 && j == 1 && rk == 0) {                                          // :262-266
        atomicAdd(cnt + hx_cell_off(W, 0, 1) + HX_SYM_GAP * HX_NSYM + a, 1u);
        sent++;
    } else if (pj == N && j - i == 1) {                                         // :271-275
        atomicAdd(cnt + hx_cell_off(W, N, N + 1) + b * HX_NSYM + HX_SYM_GAP, 1u);
        sent++;
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k1_pairs_red(const int32_t *__restrict__ rank, const int64_t *__restrict__ off,
             const uint8_t *__restrict__ codes, int64_t n_reads, int N, int W,
             uint32_t *__restrict__ cnt, unsigned long long *__restrict__ totals,
             int *__restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (BLOCK / 32);
    unsigned long long t_slices = 0, t_crumbs = 0, t_cov = 0, t_sent = 0;
    for (int64_t r = (int64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5); r < n_reads; r += nwarps) {
        const int64_t o = off[r];
        const int64_t k64 = off[r + 1] - o;
        if (k64 < 2) continue;                                                  // util.py:230
        const int rk = rank[r];
        if (rk < 0 || (int64_t)rk + k64 > N || k64 - 1 > W) {
            if (lane == 0) atomicOr(err, 1);
            continue;
        }
        const int k = (int)k64;
        const uint8_t *__restrict__ c = codes + o;
        if (lane == 0) t_slices++;
        for (int t = lane; t < k; t += 32) {
            const unsigned a = c[t];
            if (a > 6) atomicOr(err, 2);
            const bool v = (a != HX_SYM_N && a != HX_SYM_GAP && a <= 6);
            t_cov += v;                                                         // util.py:239
            t_crumbs += v ? (unsigned)(k - 1 - t) : 0u;                         // pairs with valid a
        }
        const int64_t npairs = (int64_t)k * (k - 1) / 2;
        const int m = 2 * k - 1;
        for (int64_t q = lane; q < npairs; q += 32) {
            // row-major upper triangle: row i starts at i*(m-i)/2
            const float disc = (float)((int64_t)m * m - 8 * q);
            int i = (int)(((float)m - sqrtf(disc)) * 0.5f);
            i = max(0, min(i, k - 2));
            while ((int64_t)(i + 1) * (m - (i + 1)) / 2 <= q) ++i;
            while ((int64_t)i * (m - i) / 2 > q) --i;
            const int j = i + 1 + (int)(q - (int64_t)i * (m - i) / 2);
            const unsigned a = c[i], b = c[j];
            if (a == HX_SYM_N || a == HX_SYM_GAP || a > 6 || b > 6) continue;   // util.py:258
            add_pair(cnt, N, W, rk, i, j, a, b, t_sent);
        }
    }
    flush_totals<BLOCK>(t_slices, t_crumbs, t_cov, t_sent, totals);
}

}  // namespace

int hx_launch_ingest(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off,
                     const uint8_t *d_codes, int64_t n_reads) {
    if (n_reads <= 0) return HX_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    constexpr int BLOCK = 256;
    int64_t want = (n_reads + (BLOCK / 32) - 1) / (BLOCK / 32);
    int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    HX_CUDA(cudaEventRecord(h->ev0, h->stream));
    k1_pairs_red<BLOCK><<<grid, BLOCK, 0, h->stream>>>(d_rank, d_off, d_codes, n_reads, h->N, h->W,
                                                       h->cnt, h->d_totals, h->d_err);
    h->launches++;
    HX_CUDA(cudaGetLastError());
    HX_CUDA(cudaEventRecord(h->ev1, h->stream));
    h->ev_rec = true;
    return HX_OK;
}
