grep -m1 "model name" /proc/cpuinfo; nproc; grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -E "avx2|avx512f|bmi2" | tr '\n' ' '; echo
for simd in 0 1 2; do for nt in 1 4 8 16; do echo -n "simd $simd "; HX_DENSE_SIMD=$simd tools/dense_bench $nt 0 10000000 | tail -1; done; done
for nt in 1 4 8 16; do tools/dense_bench $nt 1 10000000 | tail -1; done
