// Micro-benchmarks that decide the ingestion kernel design on B200:
// global RED (spread / L2-resident), shared ATOMS, POPC, VOTE(ballot), LOP3 rates.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__global__ void k_red(uint32_t* buf, uint64_t nwords, int iters, int same_line) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = tid * 2654435761ull + 12345;
    for (int i = 0; i < iters; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        uint64_t idx = same_line ? ((x >> 20) % (nwords / 32)) * 32 + (threadIdx.x & 31) : (x >> 20) % nwords;
        atomicAdd(buf + idx, 1u);
    }
}
__global__ void k_red_f4(float* buf, uint64_t nwords, int iters) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = tid * 2654435761ull + 12345;
    for (int i = 0; i < iters; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        uint64_t idx = ((x >> 20) % (nwords / 4)) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(buf + idx), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
    }
}
__global__ void k_atoms(uint32_t* out, int iters, int nwords) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(&sh[(x >> 8) % nwords], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
__global__ void k_popc(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u, b = blockIdx.x * 40503u + 1, acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (int i = 0; i < iters; ++i) {
        acc0 += __popc(a & b); a += 0x9e3779b9u;
        acc1 += __popc(a & b); b ^= a;
        acc2 += __popc(a & b); a += 0x7f4a7c15u;
        acc3 += __popc(a & b); b += acc0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}
__global__ void k_vote(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += __ballot_sync(0xffffffffu, a & 1); a = a * 1664525u + 1013904223u;
        acc ^= __ballot_sync(0xffffffffu, a & 2); a += acc;
        acc += __ballot_sync(0xffffffffu, a & 4); a ^= 0x9e3779b9u;
        acc ^= __ballot_sync(0xffffffffu, a & 8); a += 77;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_lop(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u, b = blockIdx.x * 40503u + 1, c = 0x12345678u, d = 0x9abcdef0u;
    for (int i = 0; i < iters; ++i) {
        a = (a & b) ^ c; b = (b | c) & d; c = (c ^ d) | a; d = (d & a) ^ b;
        a = (a & b) ^ c; b = (b | c) & d; c = (c ^ d) | a; d = (d & a) ^ b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d clock %d kHz\n", sms, clk);
    uint32_t* buf; uint32_t* out;
    size_t big = 2ull << 30;
    CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 0, big)); CK(cudaMalloc(&out, 64 << 20));
    const int grid = sms * 8, block = 256, iters = 2000;
    double nops = (double)grid * block * iters;
    for (uint64_t mb : {16ull, 64ull, 512ull, 2048ull}) {
        uint64_t nw = mb * (1 << 20) / 4;
        float ms = timeit([&] { k_red<<<grid, block>>>(buf, nw, iters, 0); });
        printf("RED.u32 random over %4llu MB : %.1f Gops/s  (%.3f lanes/clk/SM @1.9GHz)\n", (unsigned long long)mb, nops / ms / 1e6, nops / ms / 1e6 / sms / 1.9);
        ms = timeit([&] { k_red<<<grid, block>>>(buf, nw, iters, 1); });
        printf("RED.u32 warp=1 line, %4llu MB : %.1f Gops/s\n", (unsigned long long)mb, nops / ms / 1e6);
        ms = timeit([&] { k_red_f4<<<grid, block>>>((float*)buf, nw, iters); });
        printf("RED.v4.f32 random,   %4llu MB : %.1f Gops/s (x4 elements)\n", (unsigned long long)mb, nops / ms / 1e6);
    }
    CK(cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int nwords : {32, 1024, 16384}) {
        float ms = timeit([&] { k_atoms<<<grid, block, 65536>>>(out, iters, nwords); });
        CK(cudaGetLastError());
        printf("ATOMS.u32 random over %5d words: %.1f Gops/s (%.2f lanes/clk/SM)\n", nwords, nops / ms / 1e6, nops / ms / 1e6 / sms / 1.9);
    }
    { float ms = timeit([&] { k_popc<<<grid, block>>>(out, iters); });
      printf("POPC(+AND+ADD): %.1f Gpopc/s (%.1f lanes/clk/SM)\n", 4 * nops / ms / 1e6, 4 * nops / ms / 1e6 / sms / 1.9); }
    { float ms = timeit([&] { k_vote<<<grid, block>>>(out, iters); });
      printf("VOTE.ballot: %.1f Gvote-lanes/s (%.1f lanes/clk/SM)\n", 4 * nops / ms / 1e6, 4 * nops / ms / 1e6 / sms / 1.9); }
    { float ms = timeit([&] { k_lop<<<grid, block>>>(out, iters); });
      printf("LOP3 chain: %.1f Glop/s (%.1f lanes/clk/SM, 8 ops counted/iter may fuse)\n", 8 * nops / ms / 1e6, 8 * nops / ms / 1e6 / sms / 1.9); }
    return 0;
}
