// log10(x) and 10**y evaluated exactly as the C library of this image does (Ubuntu GLIBC 2.39, x86-64, the FMA
// variants its ifunc resolvers pick on any CPU with FMA + AVX2) - the functions the reference reaches through
// math.log10 and float.__pow__ at gretel/gretel.py:166-187 and inside hanselx.  When two candidate alleles tie in
// real arithmetic, which one the reference picks hangs on the last bit of these two functions, so the recovery kernels
// do not call CUDA's log10 / pow (which differ from glibc's in the last place now and then) but this transcription:
// the operation sequence was read off the disassembly of libm.so.6 (log10 -> __ieee754_log10 at 0x2b6e0, which calls
// __log_fma at 0x79d50; pow -> __pow_fma at 0x7a1e0), fused multiply-adds included, and the tables are the library's
// own (glibc_tables.h, tools/glibc_tables.py).  tests/test_gpu_recover.py checks both functions bit for bit against
// the host's libm on millions of arguments; tools/glibc_math_check.cu does the same for the host build of this header.
//
// Every * and + below is a separately rounded IEEE operation (the library is built with -fmad=false; the host check
// with -ffp-contract=off); fma() is the fused one.  Inputs outside what the recovery needs (negative, NaN, infinite
// arguments) are handled as glibc does where that is one line, and documented where not.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define HX_GL_FN __host__ __device__ __forceinline__
// the tables exist twice under nvcc: in device memory and as plain host arrays (host-side checks of this header)
#define HX_GLIBC_TABLE(name, n) static __device__ const uint64_t name##_dev[n]
#include "glibc_tables.h"
#undef HX_GLIBC_TABLE
#else
#define HX_GL_FN static inline
#endif
#define HX_GLIBC_TABLE(name, n) static const uint64_t name##_host[n]
#include "glibc_tables.h"
#undef HX_GLIBC_TABLE

#if defined(__CUDA_ARCH__)
#define HX_GL_T(name) name##_dev
#else
#define HX_GL_T(name) name##_host
#endif

HX_GL_FN double hx_gl_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
HX_GL_FN uint64_t hx_gl_u64(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}

// __log_fma (sysdeps/ieee754/dbl-64/e_log.c built with -mfma -mavx2), x positive, finite and normal
HX_GL_FN double hx_gl_log(double x) {
    const uint64_t *H = HX_GL_T(hx_gl_log_hdr);
    const uint64_t ix = hx_gl_u64(x);
    // 1 - 2^-4 <= x < 1 + 0x1.09p-4: the polynomial around 1
    if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double B0 = hx_gl_f64(H[7]), B1 = hx_gl_f64(H[8]), B2 = hx_gl_f64(H[9]), B3 = hx_gl_f64(H[10]),
                     B4 = hx_gl_f64(H[11]), B5 = hx_gl_f64(H[12]), B6 = hx_gl_f64(H[13]), B7 = hx_gl_f64(H[14]),
                     B8 = hx_gl_f64(H[15]), B9 = hx_gl_f64(H[16]), B10 = hx_gl_f64(H[17]);
        const double r = x - 1.0;
        double p2 = fma(r, B2, B1);
        double p3 = fma(r, B5, B4);
        const double r2 = r * r;
        const double p5 = fma(r, B8, B7);
        p2 = fma(r2, B3, p2);
        p3 = fma(r2, B6, p3);
        const double r3 = r * r2;
        double p1 = fma(r2, B9, p5);
        p1 = fma(r3, B10, p1);
        p1 = fma(p1, r3, p3);
        p1 = fma(p1, r3, p2);
        const double two27 = 134217728.0;
        const double t = fma(r, two27, r);                 // r + r*2^27
        const double rhi = fma(-two27, r, t);              // ... - r*2^27
        const double rhi2 = rhi * rhi;
        const double rlo = r - rhi;
        const double hi = fma(rhi2, B0, r);
        const double d = r - hi;
        const double s = r + rhi;
        const double lo = fma(rhi2, B0, d);
        const double q = B0 * rlo;
        const double lo2 = fma(q, s, lo);
        const double y = fma(p1, r3, lo2);
        return hi + y;
    }
    const double ln2hi = hx_gl_f64(H[0]), ln2lo = hx_gl_f64(H[1]);
    const double A0 = hx_gl_f64(H[2]), A1 = hx_gl_f64(H[3]), A2 = hx_gl_f64(H[4]), A3 = hx_gl_f64(H[5]), A4 = hx_gl_f64(H[6]);
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 45) & 0x7f);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = hx_gl_f64(HX_GL_T(hx_gl_log_tab)[2 * i]), logc = hx_gl_f64(HX_GL_T(hx_gl_log_tab)[2 * i + 1]);
    const double z = hx_gl_f64(iz), kd = (double)k;
    const double w = fma(kd, ln2hi, logc);
    const double r = fma(z, invc, -1.0);
    const double p12 = fma(r, A2, A1);
    const double hi = r + w;
    const double r2 = r * r;
    double lo = w - hi;
    lo = lo + r;
    lo = fma(kd, ln2lo, lo);
    const double r3 = r * r2;
    const double p34 = fma(r, A4, A3);
    lo = fma(r2, A0, lo);
    const double p = fma(p34, r2, p12);
    const double y = fma(r3, p, lo);
    return y + hi;
}

// __ieee754_log10 (sysdeps/ieee754/dbl-64/e_log10.c, baseline build: no fused operations)
HX_GL_FN double hx_gl_log10(double x) {
    uint64_t ix = hx_gl_u64(x);
    if ((ix << 1) == 0) return -hx_gl_f64(0x7ff0000000000000ull);      // log10(+-0) = -inf
    if (ix >> 63) return hx_gl_f64(0x7ff8000000000000ull);             // log10(negative) = NaN
    if (ix >= 0x7ff0000000000000ull) return x + x;                     // inf, NaN
    int k = -1023;
    if (ix <= 0xfffffffffffffull) {                        // subnormal: scale up by 2^54
        x = x * 18014398509481984.0;
        ix = hx_gl_u64(x);
        k = -1077;
    }
    k += (int)(ix >> 52);
    const int i = (int)((unsigned)k >> 31);
    const uint64_t hx = (ix & 0xfffffffffffffull) | ((uint64_t)(0x3ff - i) << 52);
    const double y = (double)(k + i);
    const double t = y * 3.69423907715893078616e-13;      // log10_2lo
    const double l = hx_gl_log(hx_gl_f64(hx));
    const double z = l * 4.34294481903251816668e-01 + t;   // ivln10
    return z + y * 3.01029995663611771306e-01;             // log10_2hi
}

// __pow_fma (sysdeps/ieee754/dbl-64/e_pow.c built with -mfma -mavx2) for x = 10: pow(10.0, y), y finite.
HX_GL_FN double hx_gl_pow10(double y) {
    const uint64_t *H = HX_GL_T(hx_gl_powlog_hdr);
    const uint64_t iy = hx_gl_u64(y);
    const unsigned topy = (unsigned)(iy >> 52) & 0x7ff;
    if (topy - 0x3be > 0x7f) {                             // |y| < 2^-65 or >= 2^63 (or zero, inf, nan)
        if ((iy << 1) == 0) return 1.0;
        if (topy < 0x3be) return 1.0 + y;                  // x > 1
        if (topy == 0x7ff) return (iy << 12) ? y + y : ((iy >> 63) ? 0.0 : y);
        return (iy >> 63) ? 0.0 : hx_gl_f64(0x7ff0000000000000ull);      // under / overflow
    }
    // log_inline(10.0): hi + lo = log(10) to about 68 bits
    const double ln2hi = hx_gl_f64(H[0]), ln2lo = hx_gl_f64(H[1]);
    const double A0 = hx_gl_f64(H[2]), A1 = hx_gl_f64(H[3]), A2 = hx_gl_f64(H[4]), A3 = hx_gl_f64(H[5]), A4 = hx_gl_f64(H[6]),
                 A5 = hx_gl_f64(H[7]), A6 = hx_gl_f64(H[8]);
    const uint64_t ix = 0x4024000000000000ull;             // 10.0
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 0x7f);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const uint64_t *T = HX_GL_T(hx_gl_powlog_tab) + 4 * i;
    const double invc = hx_gl_f64(T[0]), logc = hx_gl_f64(T[2]), logctail = hx_gl_f64(T[3]);
    const double z = hx_gl_f64(iz), kd = (double)k;
    const double t1 = fma(kd, ln2hi, logc);
    const double lo1 = fma(kd, ln2lo, logctail);
    const double r = fma(z, invc, -1.0);
    const double ar = r * A0;
    const double p12 = fma(r, A2, A1);
    const double p34 = fma(r, A4, A3);
    const double t2 = r + t1;
    const double lo2 = (t1 - t2) + r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double lo3 = fma(ar, r, -ar2);
    const double hi = t2 + ar2;
    const double p56 = fma(r, A6, A5);
    const double lo4 = (t2 - hi) + ar2;
    const double p36 = fma(p56, ar2, p34);
    const double p = fma(ar2, p36, p12);
    double lo = lo1 + lo2;
    lo = lo + lo3;
    lo = lo + lo4;
    lo = fma(ar3, p, lo);
    const double loghi = hi + lo;
    const double loglo = (hi - loghi) + lo;
    // exp_inline(ehi, elo, 0)
    const double ehi = y * loghi;
    const double e1 = fma(loghi, y, -ehi);
    const double elo = fma(y, loglo, e1);
    const uint64_t ie = hx_gl_u64(ehi);
    unsigned abstop = (unsigned)(ie >> 52) & 0x7ff;
    if (abstop - 0x3c9 > 0x3e) {
        if ((int)(abstop - 0x3c9) < 0) return 1.0 + ehi;                   // |ehi| < 2^-54
        if (abstop > 0x408) return (ie >> 63) ? 0.0 : hx_gl_f64(0x7ff0000000000000ull);
        abstop = 0;                                                        // large: the careful scaling below
    }
    const uint64_t *E = HX_GL_T(hx_gl_exp_hdr);
    const double invln2N = hx_gl_f64(E[0]), shift = hx_gl_f64(E[1]), negln2hiN = hx_gl_f64(E[2]), negln2loN = hx_gl_f64(E[3]);
    const double C2 = hx_gl_f64(E[4]), C3 = hx_gl_f64(E[5]), C4 = hx_gl_f64(E[6]), C5 = hx_gl_f64(E[7]);
    double kz = fma(ehi, invln2N, shift);
    const uint64_t ki = hx_gl_u64(kz);
    kz = kz - shift;
    double rr = fma(kz, negln2hiN, ehi);
    rr = fma(kz, negln2loN, rr);
    const unsigned idx = 2 * (unsigned)(ki & 0x7f);
    const uint64_t sbits = HX_GL_T(hx_gl_exp_tab)[idx + 1] + (ki << 45);
    rr = elo + rr;
    const double c23 = fma(rr, C3, C2);
    const double tr = rr + hx_gl_f64(HX_GL_T(hx_gl_exp_tab)[idx]);
    const double rr2 = rr * rr;
    const double c45 = fma(rr, C5, C4);
    const double q = fma(c23, rr2, tr);
    const double rr4 = rr2 * rr2;
    const double tmpv = fma(c45, rr4, q);
    if (abstop == 0) {                                                     // specialcase()
        if (!(ki & 0x80000000ull)) {                                       // k > 0: 2^1009 * (scale + scale*tmp)
            const double scale = hx_gl_f64(sbits - (1009ull << 52));
            return fma(scale, tmpv, scale) * hx_gl_f64(0x7f00000000000000ull);
        }
        const uint64_t sb2 = sbits + (1022ull << 52);
        const double scale = hx_gl_f64(sb2);
        const double st = tmpv * scale;
        double yv = scale + st;
        if (fabs(yv) < 1.0) {
            const double one = yv < 0.0 ? -1.0 : 1.0;
            double l = scale - yv;
            l = l + st;
            const double h = one + yv;
            double m = one - h;
            m = m + yv;
            m = m + l;
            m = m + h;
            yv = m - one;
            if (yv == 0.0) yv = hx_gl_f64(sb2 & 0x8000000000000000ull);
        }
        return yv * hx_gl_f64(0x0010000000000000ull);                      // 0x1p-1022
    }
    const double scale = hx_gl_f64(sbits);
    return fma(tmpv, scale, scale);
}
