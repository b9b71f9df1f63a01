// Internal definitions shared by the translation units of libhanselx.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "hanselx.h"

#define HX_NSYM 7
#define HX_CELL 49
#define HX_SYM_N 4
#define HX_SYM_DEL 5
#define HX_SYM_GAP 6   // '_'
#define HX_RING 4096   // lookback window of chosen symbols kept in shared memory by the walk
#define HX_MAX_L (HX_RING - 1)

#define HX_MAX_PEERS 8
// Where count increments go: the local partial band, or - with the fused multi-GPU exchange - the GPU
// that owns band row pj (rows are dealt out in contiguous blocks of rows_per), reached through NVLink
// peer memory.  Passed to the ingestion kernels by value.
struct HxCnt {
    uint32_t *local;
    uint32_t *const *peer;           // device array [world] of every rank's buffer (own entry included)
    int world, rows_per;
#ifdef __CUDACC__
    __device__ __forceinline__ uint32_t *cell(int64_t W, int64_t pi, int64_t pj) const {
        uint32_t *base = local;
        if (world > 1) base = peer[pj / rows_per];
        return base + (pj * W + (pj - pi - 1)) * 49;
    }
#endif
};

#define HX_WIRE_SETS 3
// One staging set of the dense wire format (wire.cu): raw shipped bytes + the rebuilt packed arrays
struct hx_wire_set {
    void *raw; int32_t *rank; int64_t *off; uint8_t *codes; int64_t *partials, *run_end;
    int64_t cap_raw, cap_rank, cap_off, cap_codes, cap_partials, cap_run_end;
    cudaEvent_t copied, decoded, consumed;
};

struct hx_matrix {
    int32_t N, W, device;
    int64_t band_elems;              // (N+2)*W*49
    cudaStream_t stream;
    bool own_stream;
    float *band;                     // float32 working matrix (what the Hansel surface reads)
    uint32_t *cnt;                   // integer counts being ingested (lazily allocated)
    bool cnt_fresh;                  // cnt is still all zero (nothing ingested / received since it was cleared)
    bool cnt_ipc;                    // cnt is a plain cudaMalloc allocation shared through CUDA IPC
    int64_t cnt_elems;               // allocated uint32 elements (>= band_elems; padded for the fused exchange)
    uint32_t *peer_host[HX_MAX_PEERS];   // fused exchange: every rank's cnt (own entry = local pointer)
    uint32_t **d_peer_tbl;           // the same table on the device
    int peer_world, peer_rows_per;
    void *ipc_opened[HX_MAX_PEERS];  // pointers returned by cudaIpcOpenMemHandle (to close)
    unsigned long long *d_totals;    // [8] slices, crumbs, covered, sentinels, -, -, -, -
    int *d_err;                      // ingestion error bits
    // staging for hx_ingest_host
    int32_t *s_rank; int64_t *s_off; uint8_t *s_codes;
    int64_t cap_reads, cap_codes;
    uint16_t *s_klen; uint32_t *s_codes4; int64_t *s_scan;     // compact wire format staging
    int64_t cap_klen, cap_codes4, cap_scan;
    hx_wire_set wire[HX_WIRE_SETS]; int wire_next;   // dense wire format: rotating staging sets
    cudaStream_t copy_stream, decode_stream; // host->device copies / decode kernels of the dense format
    uint8_t *pin[2]; int64_t pin_cap[2]; cudaEvent_t pin_ev[2]; bool pin_busy[2];   // pinned encode buffers of hx_ingest_host
    // recovery scratch
    double *scnt;                    // (N+2)*8 per-site counts + total
    int32_t *vseen;                  // (N+2) valid symbols seen per site
    bool counts_dirty;
    uint8_t *d_path;                 // path buffer(s)
    int64_t cap_path;
    double *d_stats;                 // per-iteration stats
    int64_t cap_stats;
    double *d_site;                  // 2 x 3*(N+2) per-site log10 marginal (cur), (orig), marginal (alternating haplotypes)
    cudaStream_t sum_stream;         // the ordered log10 sums of a haplotype run here, off the critical path
    cudaEvent_t sum_done[2], site_ready;
    bool sum_busy[2];
    double *d_terms;                 // walk tables: (N+2)*Lw*49 log10 lookback terms + (N+2)*8 log10 marginals
    int64_t cap_terms;
    uint8_t *d_spec;                 // speculative block paths, majority-allele guess and block flags of the walk
    int64_t cap_spec;
    double *d_partials;              // block partials of the reweight reduction
    int64_t cap_partials;
    int *d_flags;                    // [0] hole site / abort flag, [1..] misc
    int64_t *d_run_end;              // (N+1) end (exclusive) of the run of reads with each rank
    int64_t *d_run_list;             // compact run list for the tensor-core kernel: ranks (int32), stops (int64), count
    void *d_jobs;                    // per-CTA job tables of the tensor-core kernel (ingest_umma.cu)
    int64_t cap_jobs;
    double *d_misc;                  // small outputs (weights etc.)
    uint32_t *d_pack;                // packed counts for the cross-GPU exchange (api.cu)
    int64_t cap_pack;
    void *h_pinned;                  // small host buffer for D2H of scalars
    int ingest_kernel;
    int ingest_sms;                  // SMs the persistent ingestion kernels may use (0 = all)
    void *lr_scratch;                // long-read ingestion scratch (ingest_long.cu)
    void *l2_scratch;                // long-read tensor-core ingestion scratch (ingest_lumma.cu)
    cudaEvent_t ev0, ev1;
    cudaEvent_t host_ev;             // orders the chunked host->device copies of hx_ingest_host
    bool ev_rec;                     // ev0/ev1 have been recorded at least once
    bool prepass_ran;                // d_flags[4] holds the sortedness verdict of the last ingestion launch
    float last_ms[3];
    int64_t launches;
};

void hx_set_error(const char *fmt, ...);

#define HX_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            hx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return HX_E_CUDA;                                                           \
        }                                                                               \
    } while (0)

#define HX_CHECK_ARG(cond)                                                              \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            hx_set_error("%s:%d argument check failed: %s", __FILE__, __LINE__, #cond); \
            return HX_E_ARG;                                                            \
        }                                                                               \
    } while (0)

// Fills run as kernels, not cudaMemsetAsync: a memset may be queued on a copy engine behind the multi-megabyte
// host->device copies of the overlapped ingestion (measured: +0.08 ms per ingestion launch while a copy is in flight).
#ifdef __CUDACC__
static __global__ void k_fill_bytes(uint8_t *p, unsigned v, size_t bytes) {
    if (reinterpret_cast<uintptr_t>(p) & 15) {        // unaligned (small) fills: plain byte stores
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x)
            p[i] = (uint8_t)v;
        return;
    }
    const size_t n16 = bytes / 16;
    const uint4 w = make_uint4(v, v, v, v);
    uint4 *p16 = reinterpret_cast<uint4 *>(p);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p16[i] = w;
    if (blockIdx.x == 0)
        for (size_t i = n16 * 16 + threadIdx.x; i < bytes; i += blockDim.x) p[i] = (uint8_t)v;
}
static inline cudaError_t hx_fill_async(void *p, int byte, size_t bytes, cudaStream_t st) {
    if (!bytes) return cudaSuccess;
    const unsigned b = (unsigned)byte & 0xffu, v = b * 0x01010101u;
    const size_t want = (bytes / 16 + 255) / 256;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    k_fill_bytes<<<grid, 256, 0, st>>>(static_cast<uint8_t *>(p), v, bytes);
    return cudaGetLastError();
}
#endif

__host__ __device__ __forceinline__ int64_t hx_cell_off(int64_t W, int64_t pi, int64_t pj) {
    return (pj * W + (pj - pi - 1)) * HX_CELL;
}

// ingest.cu
int hx_launch_ingest(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off,
                     const uint8_t *d_codes, int64_t n_reads);
int hx_launch_ingest_presorted(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off,
                               const uint8_t *d_codes, int64_t n_reads, int64_t *run_end, const int *ok);
// ingest_umma.cu
bool hx_umma_possible(const hx_matrix *h);
int hx_launch_ingest_umma(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off, const uint8_t *d_codes,
                          int64_t n_reads, const int64_t *run_end, const int *sorted_flag);
// api.cu
HxCnt hx_cnt_ref(const hx_matrix *h);
int hx_ensure_counts_buffer(hx_matrix *h);
// wire.cu
void hx_wire_free(hx_matrix *h);
void hx_wire_trace_dump();
// set once an ingestion of this process met reads that were not sorted by rank (hx_ingest_totals reads the pre-pass
// verdict back; hx_ingest_host knows from its own pass): later short-read launches queue the counting-sort tensor-core
// kernel (ingest_lumma.cu) as their unsorted-input fallback instead of one RED per pair
bool hx_unsorted_seen();
void hx_note_unsorted();
int hx_ingest_host_pipelined(hx_matrix *h, const int32_t *rank, const int64_t *off, const uint8_t *codes, int64_t n_reads,
                             bool slim, int64_t *done_reads);
// ingest_long.cu
int hx_launch_ingest_long(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off,
                          const uint8_t *d_codes, int64_t n_reads, const int *sorted_flag);
void hx_lr_free(hx_matrix *h);
// ingest_lumma.cu
int hx_launch_ingest_lumma(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off, const uint8_t *d_codes,
                           int64_t n_reads, const int *go);
int64_t hx_lumma_scratch_bytes(const hx_matrix *h, int64_t n_reads);
void hx_l2_free(hx_matrix *h);
// recover.cu
int hx_ensure_counts(hx_matrix *h);
