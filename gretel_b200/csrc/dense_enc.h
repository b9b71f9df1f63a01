// CPU encoder of the dense wire format (bampack.cpp) in steps, so that the caller can place the blob in pinned
// memory: hx_dense_begin (sizes of the fixed sections), hx_dense_pack (ONE pass over the reads: rank deltas, SNP
// counts, 2-bit alleles straight into the blob; the two exception lists into per-thread vectors),
// hx_dense_finish (the lists behind the fixed sections).
#pragma once
#include <stdint.h>

#include <vector>

#define HX_DENSE_MAX_THREADS 64

struct HxDensePlan {
    int nt, klen_bytes;
    bool slim;                  // only rank deltas + SNP counts are encoded; the allele bytes travel as they are
    int64_t n_reads, n_codes, c0, n_esc, n_exc;
    int64_t head_bytes;         // rank deltas + SNP counts + 2-bit alleles: what hx_dense_pack writes
    int64_t bytes;              // whole blob, known after hx_dense_pack
    int64_t o_klen, o_codes2, o_exc, o_esc_idx, o_esc_delta;
    std::vector<uint32_t> exc[HX_DENSE_MAX_THREADS];
    std::vector<int64_t> esc_idx[HX_DENSE_MAX_THREADS];
    std::vector<int32_t> esc_delta[HX_DENSE_MAX_THREADS];
};

// kmax_hint: an upper bound of the SNPs per read if the caller has one (band width + 1), else 0 (one pass over off)
int hx_dense_begin(const int64_t *off, int64_t n_reads, int n_threads, int64_t kmax_hint, HxDensePlan *plan, bool slim = false);
// HX_OK, HX_E_STATE (reads not sorted by rank: use the packed arrays as they are) or HX_E_ARG
int hx_dense_pack(const int32_t *rank, const int64_t *off, const uint8_t *codes, HxDensePlan *plan, uint8_t *blob);
void hx_dense_finish(const HxDensePlan *plan, uint8_t *blob);
