// K1 on the 5th-generation tensor cores: pair expansion as an integer rank-k update.
// Replaces gretel/util.py:226-286 (+ Hansel.add_observation) for rank-sorted reads of at most 32 SNPs.
//
// Reads that share a rank r (same first SNP) cover the same sites r+1..r+k.  Write such a read as a one-hot
// row x over the columns (site t, allele a in ACGT): x[4t+a] = 1.  The counts the read adds to every site pair
// (t1 < t2) and allele pair (a, b) are exactly the entries of the outer product x^T x, so a run of reads of one
// rank contributes C = X^T X with X = [reads x columns] in {0,1}.  That is an int8 GEMM with int32 accumulation
// (tcgen05.mma kind::i8: exact by construction), M = N = columns (<= 128), K = reads:
//
//   expander warps   lane = read: load its allele bytes, turn four sites at a time into four one-hot words
//                    (1 << 8*code; N, -, _ and positions past the read's end give 0) and store them with one
//                    16-byte shared-memory store.  The row of a read IS the MN-major UMMA operand layout
//                    (core matrix = 16 columns x 8 reads), so there is no transpose anywhere.
//   MMA thread       one tcgen05.mma (M=128, N=16*ceil(k/4), K=32) per 32 reads, A and B descriptors pointing at
//                    the same shared-memory tile; accumulators in TMEM (two stages of 128 columns), released
//                    to the readout warps with tcgen05.commit at the end of a run.
//   readout warps    tcgen05.ld the upper triangle of C, add it into the CTA's sliding shared-memory tile of
//                    band rows (as k1_bitsliced does), flush retired rows to HBM with integer reductions.
//
// Reads holding N, - or _, the sentinels and the totals are handled per read by the expander warps exactly as
// in the bit-sliced kernels (ingest_common.cuh).
#include "hx_internal.cuh"
#include "ingest_common.cuh"

namespace {

constexpr int UM_NSTAGE = 4;                 // X tiles in flight
constexpr int UM_XBYTES = 32768;             // bytes per X tile
constexpr int UM_EXP_WARPS = 16;             // expander warps
constexpr int UM_RD_WARPS = 4;               // readout warps = TMEM lane quarters (warp id % 4)
constexpr int UM_MMA_WARP = UM_RD_WARPS;     // warp 4 allocates TMEM and issues the MMAs
constexpr int UM_THREADS = (UM_RD_WARPS + 1 + UM_EXP_WARPS) * 32;
constexpr int UM_TMEM_COLS = 256;            // two accumulator stages x 128 int32 columns
constexpr int UM_MAX_GROUPS = 16;            // groups (of 32 reads) per X tile at most

struct UmLayout {
    int kmax;
    __host__ __device__ size_t tile_bytes() const { return (size_t)(kmax + 1) * (kmax - 1) * 64; }
    __host__ __device__ size_t bytes() const {
        return (size_t)UM_NSTAGE * UM_XBYTES + tile_bytes() + 1024;   // tile behind the X tiles (see um_desc)
    }
};

__device__ __forceinline__ uint32_t um_shl(uint32_t v, uint32_t amt) {      // shl.b32 clamps: amt >= 32 -> 0
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(amt));
    return r;
}

// Shared-memory matrix descriptor (tcgen05, SWIZZLE_NONE, MN-major): core matrix = 16 contiguous columns (bytes) x
// 8 reads at a 16-byte pitch; the next 16 columns are SBO = 128 bytes away, the next 8 reads LBO bytes away.
__device__ __forceinline__ uint64_t um_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
    return d;                                                 // base offset 0, layout type 0 = no swizzle
}

// Instruction descriptor, kind::i8: unsigned 8-bit A and B (both MN-major), int32 accumulators, M = 128.
__device__ __forceinline__ uint32_t um_idesc(int n_cols) {
    return (2u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n_cols >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void um_mma(uint32_t tmem_d, uint64_t desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %1, %2, p;\n\t}"
        ::"r"(tmem_d), "l"(desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void um_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void um_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void um_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The batch schedule, walked identically by every warp: runs of reads that share a rank (run_end[r] = index
// after the last read of rank r, -1 for ranks without reads), cut into X tiles of at most `cap` reads.
struct UmWalk {
    const int64_t *run_end;
    int64_t hi, cur, run_hi, wval;
    int N, r, wbase, cap;
    // batch
    int64_t bstart;
    int bn;
    bool first, last;

    __device__ __forceinline__ void init(const int64_t *re, int64_t lo_, int64_t hi_, int N_, int r0, int cap_) {
        run_end = re; hi = hi_; cur = lo_; run_hi = lo_; N = N_; r = r0; cap = cap_;
        wbase = r0 < 0 ? 0 : r0;
        load_window();
    }
    __device__ __forceinline__ void load_window() {
        const int i = wbase + (int)(threadIdx.x & 31);
        wval = i <= N ? run_end[i] : -1;
    }
    __device__ __forceinline__ bool next() {
        if (cur >= hi) return false;
        first = cur >= run_hi;
        if (first) {
            for (;;) {
                const unsigned m = __ballot_sync(0xffffffffu, wval > cur);
                if (m) {
                    const int l = __ffs(m) - 1;
                    r = wbase + l;
                    const int64_t e = __shfl_sync(0xffffffffu, wval, l);
                    run_hi = e < hi ? e : hi;
                    break;
                }
                wbase += 32;
                if (wbase > N) { cur = hi; return false; }      // reads with ranks outside [0,N]: flagged by the pre-pass
                load_window();
            }
        }
        bstart = cur;
        const int64_t left = run_hi - cur;
        bn = (int)(left < cap ? left : cap);
        cur += bn;
        last = cur >= run_hi;
        return true;
    }
};

// lane = read: expand its alleles into the one-hot operand row (CH chunks of four sites).
template <int CH>
__device__ __forceinline__ void um_expand_read(const uint8_t *__restrict__ codes, int64_t o, int kb, int kg,
                                               uint32_t rowaddr, uint32_t &rare_or, uint32_t &x0) {
    const uint8_t *__restrict__ c = codes + o;
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(c) & 3u);
    const uint32_t *__restrict__ cw = reinterpret_cast<const uint32_t *>(c - mis);
    const int nw = kb ? (int)((mis + kb + 3) >> 2) : 0;
    uint32_t wd[CH + 1];
#pragma unroll
    for (int w = 0; w <= CH; ++w) wd[w] = (w < nw && (w == 0 || 4 * (w - 1) < kg)) ? __ldg(cw + w) : 0u;
    const unsigned sh = 8u * mis;
    const int kb8 = 8 * kb;
    rare_or = 0;
    x0 = 0x04040404u;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
        if (4 * ch < kg) {                                        // warp-uniform
            const uint32_t xx = __funnelshift_r(wd[ch], wd[ch + 1], sh);
            const int c8 = min(max(kb8 - 32 * ch, 0), 32);        // valid bits of this word
            const uint32_t inval = um_shl(0xffffffffu, (uint32_t)c8);   // bytes past the read's end
            rare_or |= xx & 0xfcfcfcfcu & ~inval;
            const uint32_t x = (xx & ~inval) | (0x04040404u & inval);   // past the end -> 'N' -> no column
            if (ch == 0) x0 = x;
            const uint32_t y = x << 3;                            // byte j = 8 * code of site 4ch+j
            o0 = um_shl(1u, y & 0xffu);
            o1 = um_shl(1u, (y >> 8) & 0xffu);
            o2 = um_shl(1u, (y >> 16) & 0xffu);
            o3 = um_shl(1u, y >> 24);
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)ch * 128u), "r"(o0), "r"(o1),
                     "r"(o2), "r"(o3)
                     : "memory");
    }
}

template <int CH, bool FUSED>
__global__ void __launch_bounds__(UM_THREADS, 1)
k1_umma(const int32_t *__restrict__ rank, const int64_t *__restrict__ off, const uint8_t *__restrict__ codes,
        int64_t n_reads, int N, int W, int kmax, const HxCnt cnt_in, unsigned long long *__restrict__ totals,
        int *__restrict__ err, const int *__restrict__ sorted_flag, const int64_t *__restrict__ run_end) {
    extern __shared__ __align__(1024) uint8_t um_smem[];
    HxCnt cnt = cnt_in;                              // single-GPU build: the peer path folds away
    if (!FUSED) { cnt.world = 1; cnt.rows_per = 1; cnt.peer = nullptr; }
    __shared__ __align__(8) unsigned long long s_full[UM_NSTAGE], s_empty[UM_NSTAGE];
    __shared__ __align__(8) unsigned long long s_acc_full[2], s_acc_empty[2], s_meta[2];
    __shared__ int s_gk[UM_NSTAGE][UM_MAX_GROUPS];
    __shared__ int s_runkg[2];
    __shared__ uint32_t s_tmem;
    if (!*sorted_flag) return;                       // the generic fallback launch takes over

    constexpr int ROWB = 16 * CH;                    // bytes of one read's operand row
    constexpr uint32_t LBO = 128u * CH;              // 8 reads further
    constexpr int CAP = UM_XBYTES / ROWB;            // reads per X tile (512 / 256)
    constexpr int GB = CAP / 32;                     // groups per X tile
    static_assert(GB <= UM_MAX_GROUPS, "s_gk too small");

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n_reads + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per;
    const int64_t hi = lo + per < n_reads ? lo + per : n_reads;
    if (lo >= hi) return;

    const int rows = kmax + 1, cells = kmax - 1;
    const uint32_t x_saddr = ws_smem_u32(um_smem);
    uint4 *const tile = reinterpret_cast<uint4 *>(um_smem + (size_t)UM_NSTAGE * UM_XBYTES);
    uint32_t *const tile32 = reinterpret_cast<uint32_t *>(tile);

    for (size_t w = threadIdx.x; w < (size_t)rows * cells * 4; w += blockDim.x) tile[w] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int b = 0; b < UM_NSTAGE; ++b) {
            ws_mbar_init(ws_smem_u32(&s_full[b]), UM_EXP_WARPS);
            ws_mbar_init(ws_smem_u32(&s_empty[b]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ws_mbar_init(ws_smem_u32(&s_acc_full[s]), 1);
            ws_mbar_init(ws_smem_u32(&s_acc_empty[s]), UM_RD_WARPS);
            ws_mbar_init(ws_smem_u32(&s_meta[s]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UM_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ws_smem_u32(&s_tmem)),
                     "r"((uint32_t)UM_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    um_fence_before();
    __syncthreads();
    um_fence_after();
    const uint32_t tmem_base = s_tmem;

    unsigned long long t_crumbs = 0;
    unsigned n_slices = 0, n_codes = 0, n_notcov = 0, n_sent = 0, n_rcrumbs = 0, errbits = 0;

    UmWalk wk;
    wk.init(run_end, lo, hi, N, rank[lo], CAP);

    if (warp > UM_MMA_WARP) {
        // ================================ expanders ==========================================
        const int e = warp - UM_MMA_WARP - 1;
        const uint32_t pf_bytes = (uint32_t)CAP * (uint32_t)(kmax > 20 ? 24 : 16);     // ~ one X tile of codes
        for (unsigned bi = 0; wk.next(); ++bi) {
            const int b = bi % UM_NSTAGE;
            const int r = wk.r;
            const int nb = (wk.bn + 31) >> 5;
            const int64_t run_stop = wk.bstart + wk.bn;
            if (e == 0 && lane == 0 && wk.cur < hi) {         // the next X tile: its offsets
                const int64_t ahead = min((int64_t)CAP + 1, hi - wk.cur + 1);
                bs_prefetch_l2(off + wk.cur, (uint32_t)(ahead * 8));
            }
            ws_mbar_wait(ws_smem_u32(&s_empty[b]), ((bi / UM_NSTAGE) & 1) ^ 1);
            // groups are dealt round-robin over the expander warps across consecutive tiles
            int g = (e - (int)((bi * (unsigned)GB) % UM_EXP_WARPS) + UM_EXP_WARPS) % UM_EXP_WARPS;
            for (; g < nb; g += UM_EXP_WARPS) {
                const int64_t idx = wk.bstart + (int64_t)g * 32 + lane;
                int64_t o = 0;
                int kb = 0;
                if (idx < run_stop) {
                    o = off[idx];
                    const int64_t k64 = off[idx + 1] - o;
                    if (k64 >= 2) {
                        if (r < 0 || (int64_t)r + k64 > N || k64 - 1 > W) errbits |= 1;
                        else kb = (int)k64;
                    }
                    // last read of the tile: the next tile's codes start right behind it
                    if (idx + 1 == run_stop && idx + 1 < hi) bs_prefetch_l2(codes + o + k64, pf_bytes);
                }
                n_slices += kb >= 2;
                n_codes += kb;
                const int kg = __reduce_max_sync(0xffffffffu, kb);
                const int ridx = g * 32 + lane;
                const uint32_t rowaddr = x_saddr + (uint32_t)b * UM_XBYTES + (uint32_t)(ridx >> 3) * LBO + (uint32_t)(ridx & 7) * 16u;
                uint32_t rare_or, x0;
                um_expand_read<CH>(codes, o, kb, kg, rowaddr, rare_or, x0);
                if (lane == 0) s_gk[b][g] = kg;
                if (kb >= 2) {
                    const uint8_t *__restrict__ c = codes + o;
                    const unsigned a0 = x0 & 0xffu;
                    if (r == 0 && sym_valid_from(a0)) {            // util.py:262-266
                        atomicAdd(cnt.cell(W, 0, 1) + HX_SYM_GAP * HX_NSYM + a0, 1u);
                        n_sent++;
                    }
                    if (r + kb == N && !(kb == 2 && r == 0)) {     // util.py:271-275
                        const unsigned ap = c[kb - 2], bl = c[kb - 1];
                        if (sym_valid_from(ap) && bl <= 6) {
                            atomicAdd(cnt.cell(W, N, N + 1) + bl * HX_NSYM + HX_SYM_GAP, 1u);
                            n_sent++;
                        }
                    }
                }
                // reads holding N, - or _: their pairs with such an allele are added right here by the warp
                unsigned rm = __ballot_sync(0xffffffffu, kb >= 2 && rare_or != 0);
                while (rm) {
                    const int src = __ffs(rm) - 1;
                    rm &= rm - 1;
                    const int64_t o2 = __shfl_sync(0xffffffffu, o, src);
                    const int k2 = __shfl_sync(0xffffffffu, kb, src);
                    bs_rare_read(codes + o2, k2, r, W, cnt, n_rcrumbs, n_notcov, errbits);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> tensor-core reads
            __syncwarp();
            if (lane == 0) ws_mbar_arrive(ws_smem_u32(&s_full[b]));
        }
    } else if (warp == UM_MMA_WARP) {
        // ================================ MMA issuer =========================================
        unsigned ri = 0;                               // run index
        int kg_run = 0;
        bool fresh = true;                             // the run's accumulator has not been written yet
        for (unsigned bi = 0; wk.next(); ++bi) {
            const int b = bi % UM_NSTAGE;
            const int s = ri & 1;
            const int nb = (wk.bn + 31) >> 5;
            if (wk.first) {
                ws_mbar_wait(ws_smem_u32(&s_acc_empty[s]), ((ri >> 1) & 1) ^ 1);
                kg_run = 0;
                fresh = true;
            }
            ws_mbar_wait(ws_smem_u32(&s_full[b]), (bi / UM_NSTAGE) & 1);
            um_fence_after();
            if (lane == 0) {
                const uint32_t d_addr = tmem_base + (uint32_t)s * 128u;
                for (int g = 0; g < nb; ++g) {
                    const int kg = s_gk[b][g];
                    kg_run = max(kg_run, kg);
                    // the run's first MMA overwrites all 128 columns; later ones touch what their reads reach
                    const int ncols = fresh ? 16 * CH : 16 * ((max(kg, 1) + 3) >> 2);
                    const uint64_t desc = um_desc(x_saddr + (uint32_t)b * UM_XBYTES + (uint32_t)g * 4u * LBO, LBO);
                    um_mma(d_addr, desc, um_idesc(ncols), fresh ? 0u : 1u);
                    fresh = false;
                }
                um_commit(ws_smem_u32(&s_empty[b]));
                if (wk.last) {
                    s_runkg[s] = kg_run;
                    ws_mbar_arrive(ws_smem_u32(&s_meta[s]));
                    um_commit(ws_smem_u32(&s_acc_full[s]));
                }
            }
            __syncwarp();
            if (wk.last) ++ri;
        }
    } else {
        // ================================ readout ============================================
        const int npt = UM_RD_WARPS * 32;
        const int t1 = threadIdx.x >> 2, a = threadIdx.x & 3;
        const int per_row = cells * 16;
        int64_t flushed_upto = (int64_t)rank[lo] + 1;
        int rbase = 0;
        unsigned ri = 0;
        while (wk.next()) {
            const int r = wk.r;
            if (wk.first) {
                ws_pair_barrier(npt);                // the previous run's tile adds are complete
                if ((int64_t)r + 1 > flushed_upto) {
                    const int64_t lastrow = min((int64_t)r + 1, flushed_upto + rows - 2);
                    // rows pj <= r+1 can no longer be touched by this CTA
                    unsigned long long sum = 0;
                    int row = (int)((flushed_upto + 1) % rows);
                    for (int64_t pj = flushed_upto + 1; pj <= lastrow; ++pj) {
                        uint32_t *base = tile32 + (size_t)row * per_row;
                        for (int w = threadIdx.x; w < per_row; w += npt) {
                            const uint32_t v = base[w];
                            if (v) {
                                const int d = (w >> 4) + 1, ab = w & 15;
                                atomicAdd(cnt.cell(W, pj - d, pj) + (ab >> 2) * HX_NSYM + (ab & 3), v);
                                base[w] = 0;
                                sum += v;
                            }
                        }
                        if (++row == rows) row = 0;
                    }
                    t_crumbs += sum;
                    flushed_upto = (int64_t)r + 1;
                    ws_pair_barrier(npt);            // retired ring slots may be reused by this run's tile add
                }
                rbase = (int)(((int64_t)r + 1) % rows);
            }
            if (!wk.last) continue;
            const int s = ri & 1;
            const uint32_t par = (ri >> 1) & 1;
            ws_mbar_wait(ws_smem_u32(&s_meta[s]), par);
            const int kg = s_runkg[s];
            ws_mbar_wait(ws_smem_u32(&s_acc_full[s]), par);
            um_fence_after();
            const int ncols = 4 * kg;
            if (kg >= 2 && warp * 32 < ncols - 4) {                     // this lane quarter holds live rows
                const uint32_t taddr = tmem_base + (uint32_t)s * 128u + ((uint32_t)(warp * 32) << 16);
                for (int cb = warp * 32; cb < ncols; cb += 16) {
                    uint32_t v[16];
                    um_ld16(taddr + (uint32_t)cb, v);
                    um_wait_ld();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int t2 = (cb >> 2) + q;
                        if (t2 > t1 && t2 < kg) {
                            int row = rbase + t2;
                            if (row >= rows) row -= rows;
                            uint4 *cell = tile + ((size_t)row * cells + (t2 - t1 - 1)) * 4 + a;
                            uint4 c = *cell;
                            c.x += v[4 * q + 0]; c.y += v[4 * q + 1]; c.z += v[4 * q + 2]; c.w += v[4 * q + 3];
                            *cell = c;
                        }
                    }
                }
            }
            um_fence_before();
            __syncwarp();
            if (lane == 0) ws_mbar_arrive(ws_smem_u32(&s_acc_empty[s]));
            ++ri;
        }
        ws_pair_barrier(npt);
        {
            unsigned long long sum = 0;
            int row = (int)((flushed_upto + 1) % rows);
            for (int64_t pj = flushed_upto + 1; pj <= flushed_upto + rows - 1; ++pj) {
                uint32_t *base = tile32 + (size_t)row * per_row;
                for (int w = threadIdx.x; w < per_row; w += npt) {
                    const uint32_t v = base[w];
                    if (v) {
                        const int d = (w >> 4) + 1, ab = w & 15;
                        atomicAdd(cnt.cell(W, pj - d, pj) + (ab >> 2) * HX_NSYM + (ab & 3), v);
                        sum += v;
                    }
                }
                if (++row == rows) row = 0;
            }
            t_crumbs += sum;
        }
    }
    if (errbits) atomicOr(err, (int)errbits);
    um_fence_before();
    flush_totals(n_slices, t_crumbs + n_rcrumbs, (unsigned long long)n_codes - n_notcov, n_sent, totals);   // barriers inside
    if (warp == UM_MMA_WARP) {
        um_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)UM_TMEM_COLS)
                     : "memory");
    }
}

}  // namespace

bool hx_umma_possible(const hx_matrix *h) {
    const int kmax = h->W + 1;
    return kmax >= 2 && kmax <= 32;
}

// rank-sorted reads of at most 32 SNPs; run_end must hold -1 for ranks without reads (hx_launch_ingest fills it)
int hx_launch_ingest_umma(hx_matrix *h, const int32_t *d_rank, const int64_t *d_off, const uint8_t *d_codes,
                          int64_t n_reads, const int64_t *run_end, const int *sorted_flag) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int kmax = h->W + 1;
    const bool fused = h->peer_world > 1;
    const size_t smem = UmLayout{kmax}.bytes();
    int64_t grid = sms;
    const int64_t max_useful = (n_reads + 255) / 256;          // no thinner than 256 reads per CTA
    if (grid > max_useful) grid = max_useful;
#define HX_UM_LAUNCH(CH_)                                                                                       \
    do {                                                                                                        \
        auto kern = fused ? k1_umma<CH_, true> : k1_umma<CH_, false>;                                           \
        HX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        kern<<<(unsigned)grid, UM_THREADS, smem, h->stream>>>(d_rank, d_off, d_codes, n_reads, h->N, h->W, kmax, \
                                                              hx_cnt_ref(h), h->d_totals, h->d_err, sorted_flag, \
                                                              run_end);                                         \
    } while (0)
    if (kmax <= 16) HX_UM_LAUNCH(4);
    else HX_UM_LAUNCH(8);
#undef HX_UM_LAUNCH
    h->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
}
