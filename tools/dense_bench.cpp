// Host-side microbenchmark of the dense wire-format encoder (hx_dense_pack): threads, slim, reads.
// g++ -O2 -std=c++17 -Igretel_b200/csrc -Iinclude tools/dense_bench.cpp -o tools/dense_bench -Lgretel_b200 -lhanselx -Wl,-rpath,'$ORIGIN/../gretel_b200'
#include "dense_enc.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
int main(int argc, char **argv) {
    int nt = argc > 1 ? atoi(argv[1]) : 8; bool slim = argc > 2 && atoi(argv[2]);
    int64_t R = argc > 3 ? atoll(argv[3]) : 2000000; std::mt19937_64 g(1);
    std::vector<int32_t> rank(R); std::vector<int64_t> off(R + 1); off[0] = 0;
    for (int64_t r = 0; r < R; ++r) { rank[r] = (int32_t)(r * 9970 / R); off[r + 1] = off[r] + 2 + g() % 27; }
    std::vector<uint8_t> codes(off[R] + 64);
    for (auto &c : codes) { c = g() & 3; if (g() % 500 == 0) c = 4 + g() % 3; }
    HxDensePlan P;
    std::vector<uint8_t> blob;
    for (int it = 0; it < 5; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        hx_dense_begin(off.data(), R, nt, 30, &P, slim);
        if (blob.size() < (size_t)P.head_bytes + 16) blob.resize(P.head_bytes + 16);
        auto t1 = std::chrono::steady_clock::now();
        int rc = hx_dense_pack(rank.data(), off.data(), codes.data(), &P, blob.data());
        auto t2 = std::chrono::steady_clock::now();
        printf("nt %d slim %d rc %d begin %.3f ms pack %.3f ms  (%.2f GB/s codes) exc %lld\n", nt, slim, rc,
               std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(),
               off[R] / std::chrono::duration<double>(t2 - t1).count() / 1e9, (long long)P.n_exc);
    }
}
