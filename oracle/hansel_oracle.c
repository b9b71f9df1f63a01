/* CPU oracle (C restatement) for the Gretel/Hansel hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product never links or calls it.
 *
 * This is the fast twin of oracle/hansel_oracle.py (which is the literal, dense
 * restatement); tests/test_oracle_c.py proves the two agree bit-for-bit on small
 * inputs, and the Python one is pinned to the reference's own golden vectors
 * (tests/test_test.py:33-52 under /root/reference).  Recovery arithmetic is
 * PARITY UNPINNED (hanselx==0.0.92 is not vendored; see the .py header).
 *
 * Storage here is the banded layout the CUDA path uses, so full-size checks fit in
 * memory:  cell (pi,pj), 1 <= pj-pi <= W, lives at band[(pj*W + (pj-pi-1))*49 + a*7 + b].
 *
 * Follows (paths under /root/reference):
 *   or_ingest          gretel/util.py:226-286, 329-333
 *   or_counts_all      call sites gretel/cmd.py:86-92,127-143
 *   or_generate_path   gretel/gretel.py:136-189 + Hansel.get_edge_weights_at /
 *                      get_marginal_of_at (call sites gretel.py:155,182,186)
 *   or_reweight_path   gretel/gretel.py:79-98 + Hansel.reweight_observation
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NSYM 7
#define CELL 49
#define SYM_N 4
#define SYM_GAP 6   /* '_' */

static inline size_t cell_off(int64_t W, int64_t pi, int64_t pj) {
    return (size_t)((pj * W + (pj - pi - 1)) * CELL);
}

/* util.py:226-286.  totals = {slices, crumbs, covered, sentinel increments}.
 * Returns 0, or -1 if a read leaves [0,N] or needs a wider band than W. */
int or_ingest(const int32_t *rank, const int64_t *off, const uint8_t *codes, int64_t n_reads,
              int32_t N, int32_t W, uint32_t *band, int64_t *totals) {
    int64_t slices = 0, crumbs = 0, covered = 0, sent = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        const uint8_t *c = codes + off[r];
        int64_t k = off[r + 1] - off[r];
        if (k < 2) continue;                               /* util.py:230 */
        int64_t rk = rank[r];
        if (rk < 0 || rk + k > N || k - 1 > W) return -1;
        slices++;
        for (int64_t t = 0; t < k; ++t) covered += (c[t] != SYM_N && c[t] != SYM_GAP);   /* :239 */
        for (int64_t i = 0; i < k; ++i) {
            uint8_t a = c[i];
            if (a == SYM_GAP || a == SYM_N) continue;      /* :258 */
            for (int64_t j = i + 1; j < k; ++j) {
                uint8_t b = c[j];
                int64_t pi = rk + i + 1, pj = rk + j + 1;
                band[cell_off(W, pi, pj) + a * NSYM + b] += 1;     /* :267,274,280 */
                crumbs++;
                if (i == 0 && j == 1 && rk == 0) {         /* :262-266 */
                    band[cell_off(W, 0, 1) + SYM_GAP * NSYM + a] += 1;
                    sent++;
                } else if (j + rk + 1 == N && j - i == 1) { /* :271-275 */
                    band[cell_off(W, N, N + 1) + b * NSYM + SYM_GAP] += 1;
                    sent++;
                }
            }
        }
    }
    totals[0] = slices; totals[1] = crumbs; totals[2] = covered; totals[3] = sent;
    return 0;
}

void or_u32_to_f32(const uint32_t *src, float *dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

/* counts[s] at site p = sum_b H[s,b,p,p+1], accumulated in double in b order. */
static void counts_at(const float *band, int64_t W, int64_t p, double cnt[NSYM], double *total) {
    const float *cell = band + cell_off(W, p, p + 1);
    double tot = 0.0;
    for (int s = 0; s < NSYM; ++s) {
        double c = 0.0;
        for (int b = 0; b < NSYM; ++b) c += (double)cell[s * NSYM + b];
        cnt[s] = c;
        if (c > 0) tot += c;
    }
    *total = tot;
}

void or_counts_all(const float *band, int32_t N, int32_t W, double *out /* [(N+1)*8] */) {
    for (int64_t p = 0; p <= N; ++p) {
        double cnt[NSYM], tot;
        counts_at(band, W, p, cnt, &tot);
        for (int s = 0; s < NSYM; ++s) out[p * 8 + s] = cnt[s] > 0 ? cnt[s] : 0.0;
        out[p * 8 + 7] = tot;
    }
}

static int valid_seen(const float *band, int64_t W, int64_t p) {
    double cnt[NSYM], tot;
    counts_at(band, W, p, cnt, &tot);
    int v = 0;
    for (int s = 0; s < NSYM; ++s)
        if (s != SYM_N && s != SYM_GAP && cnt[s] > 0) v++;
    return v;
}

/* Hansel.get_edge_weights_at (UNPINNED; see .py).  weights[7] (0 where not a
 * candidate), cand mask returned; *total = sum of unnormalised weights. */
static int edge_weights_at(const float *band, int32_t N, int32_t W, int32_t L, int v_site_to,
                           int skip_unsym, int64_t snp, const uint8_t *path,
                           double w[NSYM], double *total_w) {
    double cnt[NSYM], tot;
    counts_at(band, W, snp, cnt, &tot);
    int mask = 0;
    double tw = 0.0;
    int v_to = v_site_to ? valid_seen(band, W, snp) : 0;
    for (int s = 0; s < NSYM; ++s) {
        w[s] = 0.0;
        if (skip_unsym && (s == SYM_N || s == SYM_GAP)) continue;
        if (!(cnt[s] > 0)) continue;
        double lw = log10(cnt[s] / tot);
        int64_t lmax = L < snp ? L : snp;
        for (int64_t l = 1; l <= lmax; ++l) {
            int64_t pf = snp - l;
            double obs = 0.0, sup = 0.0;
            if (l <= W) {
                const float *cell = band + cell_off(W, pf, snp);
                obs = (double)cell[path[pf] * NSYM + s];
                for (int a = 0; a < NSYM; ++a) sup += (double)cell[a * NSYM + s];
            }
            int v = v_site_to ? v_to : valid_seen(band, W, pf);
            double den = (double)v + sup;
            if (den == 0) continue;
            lw += log10((1.0 + obs) / den);
        }
        double ws = pow(10.0, lw);
        w[s] = ws;
        tw += ws;
        mask |= 1 << s;
    }
    if (tw > 0)
        for (int s = 0; s < NSYM; ++s)
            if (mask & (1 << s)) w[s] = w[s] / tw;
    *total_w = tw;
    return mask;
}

int or_edge_weights_at(const float *band, int32_t N, int32_t W, int32_t L, int v_site_to,
                       int skip_unsym, int32_t snp, const uint8_t *path, double *w, double *total_w) {
    return edge_weights_at(band, N, W, L, v_site_to, skip_unsym, snp, path, w, total_w);
}

/* gretel.py:136-189.  out = {hp_current, hp_original, min_marginal}.
 * Returns 0 ok, or the (positive) site index at which no branch exists (hole). */
int or_generate_path(const float *cur, const float *orig, int32_t N, int32_t W, int32_t L,
                     int v_site_to, int skip_unsym, uint8_t *path /* [N+1] */, double *out) {
    double rp = 0.0, rp_uw = 0.0, minm = INFINITY;
    path[0] = SYM_GAP;
    for (int64_t snp = 1; snp <= N; ++snp) {
        double w[NSYM], tw;
        int mask = edge_weights_at(cur, N, W, L, v_site_to, skip_unsym, snp, path, w, &tw);
        int next_m = -1;
        double next_v = 0.0;
        for (int s = 0; s < NSYM; ++s) {                  /* gretel.py:166-174: first max */
            if (!(mask & (1 << s))) continue;
            if (next_m < 0) { next_v = w[s]; next_m = s; }
            else if (w[s] > next_v) { next_v = w[s]; next_m = s; }
        }
        if (next_m < 0) return (int)snp;                  /* gretel.py:176-180 */
        double cnt[NSYM], tot;
        counts_at(cur, W, snp, cnt, &tot);
        double m = cnt[next_m] / tot;
        counts_at(orig, W, snp, cnt, &tot);
        double mo = (tot == 0) ? 0.0 : (cnt[next_m] > 0 ? cnt[next_m] : 0.0) / tot;
        if (m < minm) minm = m;
        rp += log10(m);
        rp_uw += log10(mo);
        path[snp] = (uint8_t)next_m;
    }
    out[0] = rp; out[1] = rp_uw; out[2] = minm;
    return 0;
}

static inline double rw(float *band, int64_t W, int a, int b, int64_t pi, int64_t pj, double ratio) {
    float *p = band + cell_off(W, pi, pj) + a * NSYM + b;
    double old = (double)*p;
    double nw = old - (ratio * old);
    *p = (float)nw;
    return old - nw;
}

/* gretel.py:79-98 in call order; cells outside the band are zero => contribute 0.0. */
double or_reweight_path(float *band, int32_t N, int32_t W, const uint8_t *path, double ratio) {
    double size = 0.0;
    for (int64_t i = 0; i <= N; ++i) {
        if (i >= N) {                                      /* gretel.py:83-85 */
            size += rw(band, W, path[i], path[0], i, i + 1, ratio);
            break;
        }
        int64_t j0 = i - W > 0 ? i - W : 0;
        for (int64_t j = j0; j < i; ++j) size += rw(band, W, path[j], path[i], j, i, ratio);
        /* j == i: diagonal cell, always 0 */
        size += rw(band, W, path[i], path[i + 1], i, i + 1, ratio);
    }
    return size;
}
