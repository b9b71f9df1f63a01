// CPU encoder of the dense wire format (bampack.cpp), in two steps so that the caller can place the blob in
// pinned memory: plan (validation + sizes), then fill.
#pragma once
#include <stdint.h>

#define HX_DENSE_MAX_THREADS 64

struct HxDensePlan {
    int nt, klen_bytes;
    int64_t n_reads, n_codes, c0, n_esc, n_exc, bytes;
    int64_t o_klen, o_codes2, o_exc, o_esc_idx, o_esc_delta;
    int64_t esc_at[HX_DENSE_MAX_THREADS], exc_at[HX_DENSE_MAX_THREADS];
};

// HX_OK, HX_E_STATE (reads not sorted by rank: use the packed arrays as they are) or HX_E_ARG
int hx_dense_plan(const int32_t *rank, const int64_t *off, const uint8_t *codes, int64_t n_reads, int n_threads,
                  HxDensePlan *plan);
void hx_dense_fill(const int32_t *rank, const int64_t *off, const uint8_t *codes, const HxDensePlan *plan, uint8_t *blob);
