// Microbenchmark: throughput of back-to-back tcgen05.mma (M=128, K=32 bytes) for the operand layouts and kinds the
// tensor-core ingestion kernel could use.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

template <int KIND>   // 0 = i8, 1 = f8f6f4
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) k_bench(int n_iter, int n_cols, int mn_major, int lbo, int sbo, int same_ab,
                                                   long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t s_tmem;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01000100u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : 1u;   // S32 / F32 accumulators
        const uint32_t idesc = (fmt << 4) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
                               ((uint32_t)(n_cols >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        for (int i = 0; i < n_iter; ++i) {
            const uint32_t oa = (uint32_t)(i & 7) * 4096u;
            const uint64_t da = make_desc(base + oa, lbo, sbo, 0);
            const uint64_t db = same_ab ? da : make_desc(base + 32768u + oa, lbo, sbo, 0);
            mma<KIND>(tm, da, db, idesc, i ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

int main() {
    long long *d_out, h[2];
    cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(k_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    cudaFuncSetAttribute(k_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    const int n_iter = 2000;
    struct Cfg { const char *name; int kind, n, mn, lbo, sbo, same; } cfgs[] = {
        {"i8  MN-major N=128 lbo1024 same A,B", 0, 128, 1, 1024, 128, 1},
        {"i8  MN-major N=80  lbo1024 same A,B", 0, 80, 1, 1024, 128, 1},
        {"i8  MN-major N=128 lbo1024 A != B  ", 0, 128, 1, 1024, 128, 0},
        {"f8  MN-major N=128 lbo1024 same A,B", 1, 128, 1, 1024, 128, 1},
        {"f8  MN-major N=80  lbo1024 same A,B", 1, 80, 1, 1024, 128, 1},
        {"i8  K-major  N=128 (lbo128 sbo256) ", 0, 128, 0, 128, 256, 1},
        {"f8  K-major  N=128 (lbo128 sbo256) ", 1, 128, 0, 128, 256, 1},
        {"f8  K-major  N=80  (lbo128 sbo256) ", 1, 80, 0, 128, 256, 1},
        {"i8  MN-major N=16  lbo1024 same A,B", 0, 16, 1, 1024, 128, 1},
        {"i8  MN-major N=256 lbo2048 A != B  ", 0, 256, 1, 2048, 128, 0},
        {"i8  K-major  N=128 A != B          ", 0, 128, 0, 128, 256, 0},
        {"i8  K-major  N=256 A != B          ", 0, 256, 0, 128, 256, 0},
        {"f8  MN-major N=16  lbo1024 same A,B", 1, 16, 1, 1024, 128, 1},
    };
    for (auto &c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            if (c.kind == 0) k_bench<0><<<148, 128, 66560>>>(n_iter, c.n, c.mn, c.lbo, c.sbo, c.same, d_out);
            else k_bench<1><<<148, 128, 66560>>>(n_iter, c.n, c.mn, c.lbo, c.sbo, c.same, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        printf("%s: issue %.1f cyc/mma, complete %.1f cyc/mma\n", c.name, (double)h[0] / n_iter, (double)h[1] / n_iter);
    }
    return 0;
}
