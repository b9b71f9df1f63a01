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

struct hx_matrix {
    int32_t N, W, device;
    int64_t band_elems;              // (N+2)*W*49
    cudaStream_t stream;
    bool own_stream;
    float *band;                     // float32 working matrix (what the Hansel surface reads)
    uint32_t *cnt;                   // integer counts being ingested (lazily allocated)
    unsigned long long *d_totals;    // [8] slices, crumbs, covered, sentinels, -, -, -, -
    int *d_err;                      // ingestion error bits
    // staging for hx_ingest_host
    int32_t *s_rank; int64_t *s_off; uint8_t *s_codes;
    int64_t cap_reads, cap_codes;
    uint16_t *s_klen; uint32_t *s_codes4; int64_t *s_scan;     // compact wire format staging
    int64_t cap_klen, cap_codes4, cap_scan;
    // recovery scratch
    double *scnt;                    // (N+2)*8 per-site counts + total
    int32_t *vseen;                  // (N+2) valid symbols seen per site
    bool counts_dirty;
    uint8_t *d_path;                 // path buffer(s)
    int64_t cap_path;
    double *d_stats;                 // per-iteration stats
    int64_t cap_stats;
    double *d_site;                  // 3*(N+2) per-site log10 marginal (cur), (orig), marginal
    double *d_terms;                 // walk tables: (N+2)*Lw*49 log10 lookback terms + (N+2)*8 log10 marginals
    int64_t cap_terms;
    double *d_partials;              // block partials of the reweight reduction
    int64_t cap_partials;
    int *d_flags;                    // [0] hole site / abort flag, [1..] misc
    int64_t *d_run_end;              // (N+1) end (exclusive) of the run of reads with each rank
    double *d_misc;                  // small outputs (weights etc.)
    void *h_pinned;                  // small pinned host buffer for D2H of scalars
    int ingest_kernel;
    void *lr_scratch;                // long-read ingestion scratch (ingest_long.cu)
    cudaEvent_t ev0, ev1;
    bool ev_rec;                     // ev0/ev1 have been recorded at least once
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

__host__ __device__ __forceinline__ int64_t hx_cell_off(int64_t W, int64_t pi, int64_t pj) {
    return (pj * W + (pj - pi - 1)) * HX_CELL;
}

// ingest.cu
int hx_launch_ingest(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off,
                     const uint8_t *d_codes, int64_t n_reads);
// ingest_long.cu
int hx_launch_ingest_long(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off,
                          const uint8_t *d_codes, int64_t n_reads);
void hx_lr_free(hx_matrix *h);
// recover.cu
int hx_ensure_counts(hx_matrix *h);
