// K2-K4: haplotype path recovery over the banded Hansel matrix, kept resident on the
// device for the whole walk (no host round trip per site).
//
// Replaces, for the reference (paths under /root/reference):
//   k_site_counts      Hansel.get_counts_at for every site   (cmd.py:85-92,123-145)
//   k_walk             gretel.py:143-187 + Hansel.get_edge_weights_at (call gretel.py:155)
//   k_path_stats/_sum  gretel.py:182-189 + Hansel.get_marginal_of_at  (calls :182,186)
//   k_reweight_path    gretel.py:79-98 + Hansel.reweight_observation  (calls :84,96)
//
// All probability arithmetic is float64 on float32-stored cells, in the same
// operation order as oracle/hansel_oracle.c, compiled with -fmad=false so that +,-,*,/
// round exactly like the CPU; only log10/pow may differ from glibc by an ulp.
#include <math.h>

#include "hx_internal.cuh"

namespace {

// ---- per-site counts -----------------------------------------------------------------
__global__ void k_site_counts(const float *__restrict__ band, int N, int W,
                              double *__restrict__ scnt, int32_t *__restrict__ vseen) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > N) return;
    const float *cell = band + hx_cell_off(W, p, p + 1);
    double tot = 0.0;
    int v = 0;
    for (int s = 0; s < HX_NSYM; ++s) {
        double c = 0.0;
        for (int b = 0; b < HX_NSYM; ++b) c += (double)cell[s * HX_NSYM + b];
        const bool pos = c > 0;
        scnt[(int64_t)p * 8 + s] = pos ? c : 0.0;
        if (pos) {
            tot += c;
            if (s != HX_SYM_N && s != HX_SYM_GAP) v++;
        }
    }
    scnt[(int64_t)p * 8 + 7] = tot;
    vseen[p] = v;
}

// ---- branch weights at one site (lanes 0..6 = candidate symbols) ----------------------
// Returns this lane's unnormalised weight; *cand_mask gets the candidate set.
__device__ __forceinline__ double edge_weight_lane(const float *__restrict__ band,
                                                   const double *__restrict__ scnt,
                                                   const int32_t *__restrict__ vseen, int W, int L,
                                                   int flags, int snp, const uint8_t *hist,
                                                   unsigned hist_mask, unsigned *cand_mask) {
    const int lane = threadIdx.x & 31;
    const int s = lane;
    double c = 0.0;
    if (lane < HX_NSYM) c = scnt[(int64_t)snp * 8 + lane];
    const double total = scnt[(int64_t)snp * 8 + 7];
    const bool skip_unsym = !(flags & HX_F_KEEP_UNSYMBOLS);
    const bool cand = lane < HX_NSYM && c > 0 && !(skip_unsym && (s == HX_SYM_N || s == HX_SYM_GAP));
    double ws = 0.0;
    if (cand) {
        double lw = log10(c / total);
        const int lmax = L < snp ? L : snp;
        const int v_to = vseen[snp];
        for (int l = 1; l <= lmax; ++l) {
            const int pf = snp - l;
            double obs = 0.0, sup = 0.0;
            if (l <= W) {
                const float *cell = band + hx_cell_off(W, pf, snp);
                const unsigned a = hist[pf & hist_mask];
                obs = (double)cell[a * HX_NSYM + s];
#pragma unroll
                for (int a2 = 0; a2 < HX_NSYM; ++a2) sup += (double)cell[a2 * HX_NSYM + s];
            }
            const int v = (flags & HX_F_VSITE_TO) ? v_to : vseen[pf];
            const double den = (double)v + sup;
            if (den != 0) lw += log10((1.0 + obs) / den);
        }
        ws = pow(10.0, lw);
    }
    *cand_mask = __ballot_sync(0xffffffffu, cand);
    return ws;
}

// Sum in symbol order, normalise, first-max argmax (gretel.py:166-174).
__device__ __forceinline__ int normalise_and_pick(double ws, unsigned cmask, double *wn_out,
                                                  double *tw_out) {
    double tw = 0.0;
#pragma unroll
    for (int s = 0; s < HX_NSYM; ++s) {
        const double v = __shfl_sync(0xffffffffu, ws, s);
        if ((cmask >> s) & 1u) tw += v;
    }
    const double wn = tw > 0 ? ws / tw : ws;
    int next = -1;
    double nv = 0.0;
#pragma unroll
    for (int s = 0; s < HX_NSYM; ++s) {
        const double v = __shfl_sync(0xffffffffu, wn, s);
        if ((cmask >> s) & 1u) {
            if (next < 0) { nv = v; next = s; }
            else if (v > nv) { nv = v; next = s; }
        }
    }
    *wn_out = wn;
    *tw_out = tw;
    return next;
}

// ---- the walk: one warp, strictly sequential over sites -------------------------------
__global__ void __launch_bounds__(32)
k_walk(const float *__restrict__ band, const double *__restrict__ scnt,
       const int32_t *__restrict__ vseen, int N, int W, int L, int flags,
       uint8_t *__restrict__ path, int *__restrict__ flagsd /* [0] hole site, [1] abort */) {
    __shared__ uint8_t ring[HX_RING];
    const int lane = threadIdx.x;
    if (flagsd[1]) return;                       // an earlier iteration of hx_recover hit a hole
    if (lane == 0) { ring[0] = HX_SYM_GAP; path[0] = HX_SYM_GAP; }
    __syncwarp();
    for (int snp = 1; snp <= N; ++snp) {
        unsigned cmask;
        const double ws = edge_weight_lane(band, scnt, vseen, W, L, flags, snp, ring, HX_RING - 1, &cmask);
        double wn, tw;
        const int next = normalise_and_pick(ws, cmask, &wn, &tw);
        if (next < 0) {                          // gretel.py:176-180
            if (lane == 0) { flagsd[0] = snp; flagsd[1] = 1; }
            return;
        }
        if (lane == 0) { ring[snp & (HX_RING - 1)] = (uint8_t)next; path[snp] = (uint8_t)next; }
        __syncwarp();
    }
    if (lane == 0) flagsd[0] = 0;
}

__global__ void __launch_bounds__(32)
k_edge_one(const float *__restrict__ band, const double *__restrict__ scnt,
           const int32_t *__restrict__ vseen, int W, int L, int flags, int snp,
           const uint8_t *__restrict__ path, double *__restrict__ out /* [7] w, [7] total, [8] mask */) {
    unsigned cmask;
    const double ws = edge_weight_lane(band, scnt, vseen, W, L, flags, snp, path, 0xffffffffu, &cmask);
    double wn, tw;
    (void)normalise_and_pick(ws, cmask, &wn, &tw);
    const int lane = threadIdx.x;
    if (lane < HX_NSYM) out[lane] = ((cmask >> lane) & 1u) ? wn : 0.0;
    if (lane == 0) { out[7] = tw; out[8] = (double)cmask; }
}

// ---- per-site marginals of the chosen path, then ordered sums -------------------------
__global__ void k_path_stats(const double *__restrict__ scnt_cur, const double *__restrict__ scnt_orig,
                             int N, const uint8_t *__restrict__ path, double *__restrict__ site,
                             const int *__restrict__ flagsd) {
    const int snp = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (flagsd[1] || snp > N) return;
    const int s = path[snp];
    const double m = scnt_cur[(int64_t)snp * 8 + s] / scnt_cur[(int64_t)snp * 8 + 7];
    const double to = scnt_orig[(int64_t)snp * 8 + 7];
    const double mo = to == 0 ? 0.0 : scnt_orig[(int64_t)snp * 8 + s] / to;
    const int64_t stride = (int64_t)N + 2;
    site[snp] = log10(m);
    site[stride + snp] = log10(mo);
    site[2 * stride + snp] = m;
}

// lanes 0,1,2 each run one strictly ordered accumulation (same order as gretel.py:185-189)
__global__ void __launch_bounds__(32)
k_path_sum(const double *__restrict__ site, int N, double min_remove, double *__restrict__ stats,
           const int *__restrict__ flagsd) {
    const int lane = threadIdx.x;
    if (flagsd[1]) { if (lane == 0) stats[5] = 0.0; return; }
    const int64_t stride = (int64_t)N + 2;
    if (lane < 2) {
        double acc = 0.0;
        const double *p = site + lane * stride;
        for (int snp = 1; snp <= N; ++snp) acc += p[snp];
        stats[lane] = acc;                                 // hp_current, hp_original
    } else if (lane == 2) {
        double mn = INFINITY;
        const double *p = site + 2 * stride;
        for (int snp = 1; snp <= N; ++snp) mn = p[snp] < mn ? p[snp] : mn;
        stats[2] = mn;                                     // min marginal
        stats[3] = mn < min_remove ? min_remove : mn;      // cmd.py:157-160
        stats[5] = 1.0;                                    // iteration completed
    }
}

// ---- reweight ---------------------------------------------------------------------------
// One thread per band cell (pj,d) on the path.  Closed form of the loop nest in
// gretel.py:79-98: pairs (p,q) with q<=N-1 once, adjacent pairs (i,i+1), i<=N-2, twice
// (two sequential roundings), (N-1,N) once, (p,N) with p<=N-2 never, (N,N+1) once with
// symbols (path[N], '_').
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_reweight_path(float *__restrict__ band, int N, int W, const uint8_t *__restrict__ path,
                const double *__restrict__ ratio_ptr, double ratio_val,
                double *__restrict__ partials, const int *__restrict__ flagsd) {
    __shared__ double sh[BLOCK / 32];
    double removed = 0.0;
    const bool dead = flagsd[1] != 0;
    const int64_t idx = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const int64_t pj = 1 + idx / W;
    const int d = 1 + (int)(idx % W);
    const int64_t pi = pj - d;
    if (!dead && pj <= (int64_t)N + 1 && pi >= 0) {
        int times;
        if (pj <= N - 1) times = d == 1 ? 2 : 1;
        else times = d == 1 ? 1 : 0;
        if (times) {
            const double ratio = ratio_ptr ? *ratio_ptr : ratio_val;
            const unsigned a = path[pi];
            const unsigned b = pj == (int64_t)N + 1 ? (unsigned)HX_SYM_GAP : (unsigned)path[pj];
            float *p = band + hx_cell_off(W, pi, pj) + a * HX_NSYM + b;
            double old = (double)*p;
            for (int t = 0; t < times; ++t) {
                const double nw = old - (ratio * old);
                const float stored = (float)nw;
                removed += old - nw;
                old = (double)stored;
            }
            *p = (float)old;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = removed;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < BLOCK / 32; ++w) t += sh[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_sum_partials(const double *__restrict__ partials, int64_t n, double *__restrict__ out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sh[w];
        *out = t;
    }
}

__global__ void k_reweight_matrix(float *__restrict__ band, int64_t n, double ratio) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double old = (double)band[i];
        band[i] = (float)(old - ratio * old);
    }
}

int ensure_buf(void **p, int64_t *cap, int64_t need_bytes) {
    if (*cap >= need_bytes) return HX_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    HX_CUDA(cudaMalloc(p, (size_t)need_bytes));
    *cap = need_bytes;
    return HX_OK;
}

int launch_reweight(hx_matrix *h, const uint8_t *d_path, const double *d_ratio, double ratio,
                    double *d_removed) {
    constexpr int BLOCK = 256;
    const int64_t cells = ((int64_t)h->N + 1) * h->W;
    const int64_t grid = (cells + BLOCK - 1) / BLOCK;
    int rc = ensure_buf((void **)&h->d_partials, &h->cap_partials, grid * (int64_t)sizeof(double));
    if (rc) return rc;
    k_reweight_path<BLOCK><<<(unsigned)grid, BLOCK, 0, h->stream>>>(h->band, h->N, h->W, d_path, d_ratio,
                                                                    ratio, h->d_partials, h->d_flags);
    k_sum_partials<<<1, 1024, 0, h->stream>>>(h->d_partials, grid, d_removed);
    h->launches += 2;
    h->counts_dirty = true;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
}

int launch_generate(hx_matrix *cur, hx_matrix *orig, int L, int flags, uint8_t *d_path,
                    double *d_stats, double min_remove) {
    int rc = hx_ensure_counts(cur);
    if (rc) return rc;
    rc = hx_ensure_counts(orig);
    if (rc) return rc;
    const int N = cur->N;
    k_walk<<<1, 32, 0, cur->stream>>>(cur->band, cur->scnt, cur->vseen, N, cur->W, L, flags, d_path,
                                      cur->d_flags);
    k_path_stats<<<(N + 255) / 256, 256, 0, cur->stream>>>(cur->scnt, orig->scnt, N, d_path, cur->d_site,
                                                           cur->d_flags);
    k_path_sum<<<1, 32, 0, cur->stream>>>(cur->d_site, N, min_remove, d_stats, cur->d_flags);
    cur->launches += 3;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
}

}  // namespace

int hx_ensure_counts(hx_matrix *h) {
    if (!h->counts_dirty) return HX_OK;
    k_site_counts<<<(h->N + 1 + 127) / 128, 128, 0, h->stream>>>(h->band, h->N, h->W, h->scnt, h->vseen);
    h->launches++;
    HX_CUDA(cudaGetLastError());
    h->counts_dirty = false;
    return HX_OK;
}

// --------------------------------------------------------------------------- C ABI
extern "C" {

int hx_counts_all(hx_matrix *h, double *out) {
    HX_CHECK_ARG(h && out);
    HX_CUDA(cudaSetDevice(h->device));
    int rc = hx_ensure_counts(h);
    if (rc) return rc;
    HX_CUDA(cudaMemcpyAsync(out, h->scnt, sizeof(double) * 8 * ((size_t)h->N + 1), cudaMemcpyDeviceToHost,
                            h->stream));
    HX_CUDA(cudaStreamSynchronize(h->stream));
    return HX_OK;
}

int hx_marginal_of_at(hx_matrix *h, int sym, int32_t pos, double *out) {
    HX_CHECK_ARG(h && out && sym >= 0 && sym < HX_NSYM && pos >= 0 && pos <= h->N);
    HX_CUDA(cudaSetDevice(h->device));
    int rc = hx_ensure_counts(h);
    if (rc) return rc;
    double row[8];
    HX_CUDA(cudaMemcpyAsync(row, h->scnt + (size_t)pos * 8, sizeof(row), cudaMemcpyDeviceToHost, h->stream));
    HX_CUDA(cudaStreamSynchronize(h->stream));
    *out = row[7] == 0 ? 0.0 : row[sym] / row[7];
    return HX_OK;
}

int hx_edge_weights_at(hx_matrix *h, int32_t snp, const uint8_t *path, int32_t L, int flags,
                       double weights[7], double *total, int *mask) {
    HX_CHECK_ARG(h && path && weights && total && mask && snp >= 1 && snp <= h->N && L >= 0);
    HX_CUDA(cudaSetDevice(h->device));
    int rc = hx_ensure_counts(h);
    if (rc) return rc;
    rc = ensure_buf((void **)&h->d_path, &h->cap_path, (int64_t)h->N + 2);
    if (rc) return rc;
    HX_CUDA(cudaMemcpyAsync(h->d_path, path, (size_t)snp, cudaMemcpyHostToDevice, h->stream));
    k_edge_one<<<1, 32, 0, h->stream>>>(h->band, h->scnt, h->vseen, h->W, L, flags, snp, h->d_path, h->d_misc);
    h->launches++;
    HX_CUDA(cudaGetLastError());
    double res[9];
    HX_CUDA(cudaMemcpyAsync(res, h->d_misc, sizeof(res), cudaMemcpyDeviceToHost, h->stream));
    HX_CUDA(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < 7; ++s) weights[s] = res[s];
    *total = res[7];
    *mask = (int)res[8];
    return HX_OK;
}

int hx_generate_path(hx_matrix *cur, hx_matrix *orig, int32_t L, int flags, uint8_t *out_path,
                     double out[3], int32_t *hole_site) {
    HX_CHECK_ARG(cur && orig && out_path && out && hole_site);
    HX_CHECK_ARG(cur->N == orig->N && cur->device == orig->device && L >= 0 && L <= HX_MAX_L);
    HX_CUDA(cudaSetDevice(cur->device));
    const int N = cur->N;
    int rc = ensure_buf((void **)&cur->d_path, &cur->cap_path, (int64_t)N + 2);
    if (rc) return rc;
    rc = ensure_buf((void **)&cur->d_stats, &cur->cap_stats, 8 * (int64_t)sizeof(double));
    if (rc) return rc;
    // orig's counts are produced on orig's stream; order them before our walk
    if (orig->stream != cur->stream && orig->counts_dirty) {
        rc = hx_ensure_counts(orig);
        if (rc) return rc;
        HX_CUDA(cudaStreamSynchronize(orig->stream));
    }
    HX_CUDA(cudaMemsetAsync(cur->d_flags, 0, 2 * sizeof(int), cur->stream));
    HX_CUDA(cudaEventRecord(cur->ev0, cur->stream));
    rc = launch_generate(cur, orig, L, flags, cur->d_path, cur->d_stats, 0.0);
    if (rc) return rc;
    HX_CUDA(cudaEventRecord(cur->ev1, cur->stream));
    cur->ev_rec = true;
    int hflags[2];
    double stats[8];
    HX_CUDA(cudaMemcpyAsync(hflags, cur->d_flags, sizeof(hflags), cudaMemcpyDeviceToHost, cur->stream));
    HX_CUDA(cudaMemcpyAsync(stats, cur->d_stats, sizeof(stats), cudaMemcpyDeviceToHost, cur->stream));
    HX_CUDA(cudaMemcpyAsync(out_path, cur->d_path, (size_t)N + 1, cudaMemcpyDeviceToHost, cur->stream));
    HX_CUDA(cudaStreamSynchronize(cur->stream));
    cudaEventElapsedTime(&cur->last_ms[1], cur->ev0, cur->ev1);
    if (hflags[1]) {
        *hole_site = hflags[0];
        return HX_HOLE;
    }
    *hole_site = 0;
    out[0] = stats[0]; out[1] = stats[1]; out[2] = stats[2];
    return HX_OK;
}

int hx_reweight_path(hx_matrix *h, const uint8_t *path, double ratio, double *removed) {
    HX_CHECK_ARG(h && path && removed);
    HX_CUDA(cudaSetDevice(h->device));
    int rc = ensure_buf((void **)&h->d_path, &h->cap_path, (int64_t)h->N + 2);
    if (rc) return rc;
    HX_CUDA(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(int), h->stream));
    HX_CUDA(cudaMemcpyAsync(h->d_path, path, (size_t)h->N + 1, cudaMemcpyHostToDevice, h->stream));
    HX_CUDA(cudaEventRecord(h->ev0, h->stream));
    rc = launch_reweight(h, h->d_path, nullptr, ratio, h->d_misc);
    if (rc) return rc;
    HX_CUDA(cudaEventRecord(h->ev1, h->stream));
    h->ev_rec = true;
    HX_CUDA(cudaMemcpyAsync(removed, h->d_misc, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    HX_CUDA(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->last_ms[2], h->ev0, h->ev1);
    return HX_OK;
}

int hx_recover(hx_matrix *cur, hx_matrix *orig, int32_t L, int flags, int32_t max_paths,
               double min_remove, uint8_t *paths, double *stats, int32_t *n_found) {
    HX_CHECK_ARG(cur && orig && paths && stats && n_found && max_paths >= 0);
    HX_CHECK_ARG(cur->N == orig->N && cur->device == orig->device && L >= 0 && L <= HX_MAX_L);
    HX_CUDA(cudaSetDevice(cur->device));
    const int64_t N = cur->N;
    *n_found = 0;
    if (max_paths == 0) return HX_OK;
    int rc = ensure_buf((void **)&cur->d_path, &cur->cap_path, (N + 1) * (int64_t)max_paths);
    if (rc) return rc;
    rc = ensure_buf((void **)&cur->d_stats, &cur->cap_stats, 8 * (int64_t)sizeof(double) * max_paths);
    if (rc) return rc;
    if (orig->stream != cur->stream && orig->counts_dirty) {
        rc = hx_ensure_counts(orig);
        if (rc) return rc;
        HX_CUDA(cudaStreamSynchronize(orig->stream));
    }
    HX_CUDA(cudaMemsetAsync(cur->d_flags, 0, 2 * sizeof(int), cur->stream));
    HX_CUDA(cudaMemsetAsync(cur->d_stats, 0, 8 * sizeof(double) * (size_t)max_paths, cur->stream));
    HX_CUDA(cudaEventRecord(cur->ev0, cur->stream));
    for (int it = 0; it < max_paths; ++it) {
        uint8_t *dp = cur->d_path + (size_t)it * (N + 1);
        double *ds = cur->d_stats + (size_t)it * 8;
        rc = launch_generate(cur, orig, L, flags, dp, ds, min_remove);
        if (rc) return rc;
        rc = launch_reweight(cur, dp, ds + 3, 0.0, ds + 4);
        if (rc) return rc;
    }
    HX_CUDA(cudaEventRecord(cur->ev1, cur->stream));
    cur->ev_rec = true;
    double *hs = (double *)malloc(sizeof(double) * 8 * (size_t)max_paths);
    if (!hs) return HX_E_NOMEM;
    cudaError_t e = cudaMemcpyAsync(hs, cur->d_stats, sizeof(double) * 8 * (size_t)max_paths,
                                    cudaMemcpyDeviceToHost, cur->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(paths, cur->d_path, (size_t)(N + 1) * max_paths, cudaMemcpyDeviceToHost, cur->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cur->stream);
    if (e != cudaSuccess) {
        free(hs);
        hx_set_error("hx_recover: %s", cudaGetErrorString(e));
        return HX_E_CUDA;
    }
    cudaEventElapsedTime(&cur->last_ms[1], cur->ev0, cur->ev1);
    int found = 0;
    for (int it = 0; it < max_paths; ++it) {
        if (hs[(size_t)it * 8 + 5] != 1.0) break;
        for (int q = 0; q < 5; ++q) stats[(size_t)it * 5 + q] = hs[(size_t)it * 8 + q];
        found++;
    }
    free(hs);
    *n_found = found;
    return HX_OK;
}

int hx_reweight_matrix(hx_matrix *h, double ratio) {
    HX_CHECK_ARG(h);
    HX_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->band_elems;
    k_reweight_matrix<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->band, n, ratio);
    h->launches++;
    h->counts_dirty = true;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
}

}  // extern "C"
