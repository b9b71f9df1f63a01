// Host check of gretel_b200/csrc/glibc_math.cuh against the C library it transcribes:
//   g++ -O2 -mfma -ffp-contract=off -I gretel_b200/csrc -o /tmp/glibc_math_check tools/glibc_math_check.cpp -lm && /tmp/glibc_math_check
#include <stdio.h>
#include <stdlib.h>
#include "glibc_math.cuh"

static uint64_t s = 88172645463325252ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double u01() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv) {
    const long n = argc > 1 ? atol(argv[1]) : 20000000;
    long bad_log = 0, bad_pow = 0;
    for (long i = 0; i < n; ++i) {
        double x;
        switch (i & 7) {
            case 0: x = u01(); break;                                   // probabilities
            case 1: x = (1.0 + (double)(rnd() % 100000)) / (1.0 + (double)(rnd() % 100000) + (double)(rnd() % 100000)); break;   // count ratios
            case 2: x = 1.0 + (u01() - 0.5) * 0.2; break;               // near 1
            case 3: x = exp((u01() - 0.5) * 1400.0); break;             // the whole range
            case 4: x = 1.0 / (double)(1 + rnd() % 64); break;          // 1/v
            case 5: x = (double)(1 + rnd() % 1000000); break;
            case 6: x = u01() * 1e-310; if (x == 0) x = 5e-324; break;  // subnormal
            default: { uint64_t b = rnd() & 0x7fefffffffffffffull; memcpy(&x, &b, 8); if (!(x > 0)) x = 1.5; }
        }
        const double a = hx_gl_log10(x), b = log10(x);
        if (memcmp(&a, &b, 8)) { if (bad_log++ < 10) printf("log10(%a): mine %a libm %a\n", x, a, b); }
        double y;
        switch (i & 3) {
            case 0: y = -u01() * 330.0; break;
            case 1: y = -u01() * 20.0; break;
            case 2: y = (u01() - 0.5) * 700.0; break;
            default: y = (u01() - 0.5) * 1e-3; break;
        }
        if ((i & 1023) == 0) y = 0.0;
        if ((i & 1023) == 1) y = -u01() * 1e-20;
        if ((i & 1023) == 2) y = -307.0 - u01() * 20.0;
        const double c = hx_gl_pow10(y), d = pow(10.0, y);
        if (memcmp(&c, &d, 8)) { if (bad_pow++ < 10) printf("pow(10, %a): mine %a libm %a\n", y, c, d); }
    }
    printf("%ld arguments each: log10 mismatches %ld, pow10 mismatches %ld\n", n, bad_log, bad_pow);
    return bad_log || bad_pow;
}
