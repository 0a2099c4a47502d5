// csrc/jmid_gemm.cu -- persistent, warp-specialised tcgen05 GEMM for the denoiser's dense layers.
//
//   C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)      A, W bf16 K-major; fp32 accumulators in TMEM
//
// Replaces the cuBLAS calls behind nn.Linear / in_proj / out_proj / linear1 / linear2 in
// sicnav_diffusion/JMID/MID/models/diffusion.py:161-171 (through torch 1.13) with one kernel family:
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles A[128x64], W[BNx64] into a 4-stage smem ring
//   warp 1      MMA issuer:   one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per stage,
//                             tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..5  epilogue:     tcgen05.ld (software pipelined) -> bias (smem) / ReLU / ConcatSquash gate -> 128B-swizzled
//                             smem staging -> cp.async.bulk.tensor store (coalesced, asynchronous, clips the M tail)
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the main loop of
// tile i+1; CTAs are persistent (grid = #SMs) and walk the tile list with N fastest so that concurrently running
// CTAs share the same A rows in L2.
#include <cstdlib>
#include <mutex>

#include "jmid_internal.h"
#include "tc_utils.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
constexpr int EPI_WARPS = 8;
constexpr int STAGING_BYTES = 32 * 128; // per epilogue warp: 32 rows x 128 B (one swizzle atom wide)

template <int BN> struct GemmCfg {
    static constexpr int STAGES = 4;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN; // 512 or 256: both powers of two
    static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_STAGING + EPI_WARPS * STAGING_BYTES;
    static constexpr int OFF_BARS = OFF_BIAS + BN * 4;
    static constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024 /*align*/;
};

// one 32-column chunk of one accumulator row: + bias (+ReLU | ConcatSquash), result left in v[]
// Per-thread (= per accumulator row) state of the LayerNorm fold (see the comment above gemm2_bf16_tn_kernel).
struct RowFold {
    float rstd, nmr;       // FOLD: the A operand was z, not LN(z):  v = rstd * acc - rstd * mu * colsum[n] + bias'[n];  nmr = -rstd * mu
    float r_rstd, r_mu;    // RES == 2: the residual is LN(z_res) recomputed from z_res and its row statistics
    uint64_t sum2, sq2;    // RES != 0: sum / sum of squares of the row this thread writes, two interleaved partial sums each (FADD2 / FFMA2)
};

// Packed-pair (f32x2) forms of the fold epilogue: one FFMA2 / FADD2 per TWO accumulator columns.  The drain of an out-proj tile
// (K = 512: 2048 clk of MMAs per tile) spent ~2500 issue slots per scheduler on scalar epilogue arithmetic and was the limiter of
// the RES kernels (ncu: neither DRAM nor the tensor pipe above 55 %); packed it is about half of that.
template <int EPI, bool FOLD>
__device__ __forceinline__ void epilogue_math2(const uint32_t (&acc)[32], uint64_t (&v2)[16], const float *s_bias, const GemmEpi &ep,
                                               int col0_global, int ba, const RowFold &rf, const float *s_colsum = nullptr)
{
    const ulonglong2 *b2 = reinterpret_cast<const ulonglong2 *>(s_bias);
    if constexpr (FOLD) {
        const ulonglong2 *c2 = reinterpret_cast<const ulonglong2 *>(s_colsum);   // shared memory (ncu: as eight LDG.128 per chunk it cost 6 % of the kernel)
        const uint64_t rstd2 = tc::f2_pack(rf.rstd, rf.rstd), nmr2 = tc::f2_pack(rf.nmr, rf.nmr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 b = b2[j], c = c2[j];
            v2[2 * j] = tc::f2_fma(tc::f2_pack(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1])), rstd2, tc::f2_fma(nmr2, c.x, b.x));
            v2[2 * j + 1] = tc::f2_fma(tc::f2_pack(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])), rstd2, tc::f2_fma(nmr2, c.y, b.y));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 b = b2[j];
            v2[2 * j] = tc::f2_add(tc::f2_pack(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1])), b.x);
            v2[2 * j + 1] = tc::f2_add(tc::f2_pack(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])), b.y);
        }
    }
    if constexpr (EPI == EPI_BIAS_RELU_BF16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float lo, hi;
            tc::f2_unpack(v2[j], lo, hi);
            v2[j] = tc::f2_pack(fmaxf(lo, 0.0f), fmaxf(hi, 0.0f));
        }
    }
    if constexpr (EPI == EPI_CSL_BF16) {
        const ulonglong2 *g2 = reinterpret_cast<const ulonglong2 *>(ep.gate + (size_t)ba * ep.tab_ld + col0_global);
        const ulonglong2 *h2 = reinterpret_cast<const ulonglong2 *>(ep.hbias + (size_t)ba * ep.tab_ld + col0_global);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 g = __ldg(g2 + j), h = __ldg(h2 + j);
            v2[2 * j] = tc::f2_fma(v2[2 * j], g.x, h.x);
            v2[2 * j + 1] = tc::f2_fma(v2[2 * j + 1], g.y, h.y);
        }
    }
}

// + residual for one 32-column chunk (`rz` = the thread's 32 residual bf16 of these columns), row statistics, bf16 pairs in pk[16].
// RES == 2: the residual is LayerNorm(z_res) = gamma (z rr + rn) + beta with rr = rstd, rn = -mu rstd of the residual's row; beta was
// added to the bias slice in shared memory when the tile started, so a pair costs two FFMA2.  The statistics are those of the ROUNDED
// values (what the consuming GEMM reads); taking them before rounding (-DSNB_STATS_UNROUNDED) saves two integer ops per pair, is not
// faster in the step and costs parity with the shipped checkpoint (eps 2.1e-2 instead of 1.9e-2).
template <int RES>
__device__ __forceinline__ void residual_pack2(uint64_t (&v2)[16], const uint4 (&rz)[4], const GemmEpi &ep, int col0_global, RowFold &rf,
                                               uint32_t (&pk)[16])
{
    if constexpr (RES != 0) {
        const uint32_t w[16] = {rz[0].x, rz[0].y, rz[0].z, rz[0].w, rz[1].x, rz[1].y, rz[1].z, rz[1].w,
                                rz[2].x, rz[2].y, rz[2].z, rz[2].w, rz[3].x, rz[3].y, rz[3].z, rz[3].w};
        if constexpr (RES == 2) {
            const ulonglong2 *g2 = reinterpret_cast<const ulonglong2 *>(ep.res_gamma + col0_global);
            const float rn = -rf.r_mu * rf.r_rstd;
            const uint64_t rr2 = tc::f2_pack(rf.r_rstd, rf.r_rstd), rn2 = tc::f2_pack(rn, rn);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const ulonglong2 g = __ldg(g2 + j);
                const uint64_t za = tc::f2_pack(__uint_as_float(w[2 * j] << 16), __uint_as_float(w[2 * j] & 0xffff0000u));
                const uint64_t zb = tc::f2_pack(__uint_as_float(w[2 * j + 1] << 16), __uint_as_float(w[2 * j + 1] & 0xffff0000u));
                v2[2 * j] = tc::f2_fma(g.x, tc::f2_fma(za, rr2, rn2), v2[2 * j]);
                v2[2 * j + 1] = tc::f2_fma(g.y, tc::f2_fma(zb, rr2, rn2), v2[2 * j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                v2[j] = tc::f2_add(v2[j], tc::f2_pack(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u)));
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float lo, hi;
        tc::f2_unpack(v2[j], lo, hi);
        pk[j] = tc::pack_bf16(lo, hi);
        if constexpr (RES != 0) {
#ifdef SNB_STATS_UNROUNDED
            const uint64_t x2 = v2[j];
#else
            const uint64_t x2 = tc::f2_pack(__uint_as_float(pk[j] << 16), __uint_as_float(pk[j] & 0xffff0000u));   // what the consumer will read
#endif
            rf.sum2 = tc::f2_add(rf.sum2, x2);
            rf.sq2 = tc::f2_fma(x2, x2, rf.sq2);
        }
    }
}

template <int EPI, bool FOLD = false>
__device__ __forceinline__ void epilogue_math(const uint32_t (&acc)[32], float (&v)[32], const float *s_bias, const GemmEpi &ep,
                                              int col0_global, int ba, const RowFold &rf = RowFold())
{
    const float4 *b4 = reinterpret_cast<const float4 *>(s_bias);
    if constexpr (FOLD) {
        const float4 *c4 = reinterpret_cast<const float4 *>(ep.colsum + col0_global);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j], c = __ldg(c4 + j);
            v[4 * j + 0] = fmaf(__uint_as_float(acc[4 * j + 0]), rf.rstd, fmaf(rf.nmr, c.x, b.x));
            v[4 * j + 1] = fmaf(__uint_as_float(acc[4 * j + 1]), rf.rstd, fmaf(rf.nmr, c.y, b.y));
            v[4 * j + 2] = fmaf(__uint_as_float(acc[4 * j + 2]), rf.rstd, fmaf(rf.nmr, c.z, b.z));
            v[4 * j + 3] = fmaf(__uint_as_float(acc[4 * j + 3]), rf.rstd, fmaf(rf.nmr, c.w, b.w));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
            v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
        }
    }
    if constexpr (EPI == EPI_BIAS_RELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    if constexpr (EPI == EPI_CSL_BF16) {
        const float4 *g4 = reinterpret_cast<const float4 *>(ep.gate + (size_t)ba * ep.tab_ld + col0_global);
        const float4 *h4 = reinterpret_cast<const float4 *>(ep.hbias + (size_t)ba * ep.tab_ld + col0_global);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 g = __ldg(g4 + j), h = __ldg(h4 + j);
            v[4 * j + 0] = fmaf(v[4 * j + 0], g.x, h.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], g.y, h.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], g.z, h.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], g.w, h.w);
        }
    }
}

// 16-byte chunk `c` (0..7) of row `r` (0..31) inside a 32 x 128 B staging tile with the TMA 128-byte swizzle
__device__ __forceinline__ uint4 *staging_slot(uint8_t *stg, int r, int c) { return reinterpret_cast<uint4 *>(stg + r * 128 + ((c ^ (r & 7)) << 4)); }

// Drains columns [c_begin, c_end) (multiples of 64) of one warp's 32 accumulator rows: software-pipelined tcgen05.ld ->
// bias / ReLU / ConcatSquash -> 128B-swizzled staging tile -> one TMA store per 128 bytes of output row.
template <int EPI>
__device__ __forceinline__ void epilogue_drain(uint32_t t_addr, uint8_t *stg, const float *s_bias, const GemmEpi &ep,
                                               const CUtensorMap *tmC, int n0, int row0, int ba, int c_begin, int c_end, int lane)
{
    constexpr bool OUT_F32 = (EPI == EPI_BIAS_F32);
    uint32_t ra[32], rb[32];
    float v[32];
    tc::tmem_ld_32x32(t_addr + c_begin, ra);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 64) {
        tc::tmem_ld_wait();
        tc::tmem_ld_32x32(t_addr + c + 32, rb);
        epilogue_math<EPI>(ra, v, s_bias + c, ep, n0 + c, ba);
        if (lane == 0) tc::tma_store_wait_read<0>();
        __syncwarp();
        if constexpr (OUT_F32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) *staging_slot(stg, lane, j) = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tc::tma_store_2d(tmC, stg, n0 + c, row0); tc::tma_store_commit(); }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *staging_slot(stg, lane, j) = make_uint4(tc::pack_bf16(v[8 * j], v[8 * j + 1]), tc::pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                         tc::pack_bf16(v[8 * j + 4], v[8 * j + 5]), tc::pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        }
        tc::tmem_ld_wait();
        if (c + 64 < c_end) tc::tmem_ld_32x32(t_addr + c + 64, ra);
        epilogue_math<EPI>(rb, v, s_bias + c + 32, ep, n0 + c + 32, ba);
        if constexpr (OUT_F32) {
            if (lane == 0) tc::tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) *staging_slot(stg, lane, j) = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tc::tma_store_2d(tmC, stg, n0 + c + 32, row0); tc::tma_store_commit(); }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *staging_slot(stg, lane, 4 + j) = make_uint4(tc::pack_bf16(v[8 * j], v[8 * j + 1]), tc::pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                             tc::pack_bf16(v[8 * j + 4], v[8 * j + 5]), tc::pack_bf16(v[8 * j + 6], v[8 * j + 7]));
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tc::tma_store_2d(tmC, stg, n0 + c, row0); tc::tma_store_commit(); }
        }
    }
}

// Drain of one warp's 32 rows x 128 columns (the pair kernel, BN = 256, bf16 output) of a CONSUMING GEMM of the LayerNorm fold
// (FOLD; the residual-adding kernels have their own drain below).  s_bias / s_colsum point into the kernel's shared-memory tables of
// ALL N columns at this chunk's first column.
template <int EPI, bool FOLD>
__device__ __forceinline__ void epilogue_drain_fold(uint32_t t_addr, uint8_t *stg, const float *s_bias, const float *s_colsum, const GemmEpi &ep,
                                                    const CUtensorMap *tmC, int n0, int row0, int ba, int c_begin, int lane, RowFold &rf)
{
    uint32_t ra[32], rb[32], pk[16];
    uint64_t v[16];
    const uint4 none[4] = {};
    tc::tmem_ld_32x32(t_addr + c_begin, ra);
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
        const int c = c_begin + ci * 64;
        tc::tmem_ld_wait();
        tc::tmem_ld_32x32(t_addr + c + 32, rb);
        epilogue_math2<EPI, FOLD>(ra, v, s_bias + c, ep, n0 + c, ba, rf, s_colsum + c);
        residual_pack2<0>(v, none, ep, n0 + c, rf, pk);
        if (lane == 0) tc::tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) *staging_slot(stg, lane, j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tc::tmem_ld_wait();
        if (ci == 0) tc::tmem_ld_32x32(t_addr + c + 64, ra);
        epilogue_math2<EPI, FOLD>(rb, v, s_bias + c + 32, ep, n0 + c + 32, ba, rf, s_colsum + c + 32);
        residual_pack2<0>(v, none, ep, n0 + c + 32, rf, pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) *staging_slot(stg, lane, 4 + j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tc::tma_store_2d(tmC, stg, n0 + c, row0); tc::tma_store_commit(); }
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmEpi ep, const int M, const int N, const int K)
{
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::OFF_BARS);
    uint64_t *full = bars, *empty = bars + Cfg::STAGES, *tfull = bars + 2 * Cfg::STAGES, *tempty = bars + 2 * Cfg::STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * Cfg::STAGES + 4);
    float *s_bias = reinterpret_cast<float *>(smem + Cfg::OFF_BIAS);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int m_tiles = (M + BM - 1) / BM, n_tiles = N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = K / BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        tc::prefetch_tmap(&tmC);
        for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], EPI_WARPS); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *a = smem + stage * Cfg::STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    tc::tma_load_2d(a, &tmA, &full[stage], kb * BK, m0);
                    tc::tma_load_2d(a + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Issue cost (tools/mma_bench*.cu): a descriptor built per tcgen05.mma from per-thread registers costs ~130 clk on the
        // issuing thread.  The warp index and the TMEM base are made provably warp-uniform (shuffle broadcasts), one elected
        // thread issues, and inside the unrolled K loop an MMA's descriptors are `stage base + compile-time constant`.
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN, 0, 0);
            const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_u + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint64_t da = da0 + (uint64_t)stage * (Cfg::STAGE_BYTES >> 4), db = da + (Cfg::A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc::umma_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc::umma_commit(&empty[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int quarter = warp & 3;            // TMEM lanes [32*quarter, 32*quarter + 32)
        const int half = (warp - 2) >> 2;        // which half of the tile's BN columns this warp drains
        const int etid = (warp - 2) * 32 + lane;
        uint8_t *stg = smem + Cfg::OFF_STAGING + (warp - 2) * STAGING_BYTES;
        int it = 0, bias_n0 = -1;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
            const int row0 = m0 + quarter * 32;
            int ba = 0;
            if constexpr (EPI == EPI_CSL_BF16) {
                int row = row0 + lane;
                if (row >= M) row = M - 1;
                const int b = row / ep.tok_per_env;
                const int r = (row - b * ep.tok_per_env) / ep.T;
                ba = b * ep.A + (r % ep.A);
            }
            // this tile's bias slice -> smem, only when the column block changed (concat4, N = 128, has ONE column block: filled once);
            // the previous tile's readers are past their last use: barrier below
            if (n0 != bias_n0) {
                tc::named_bar_sync(1, EPI_WARPS * 32);
                for (int i = etid; i < BN; i += EPI_WARPS * 32) s_bias[i] = __ldg(ep.bias + n0 + i);
                tc::named_bar_sync(1, EPI_WARPS * 32);
                bias_n0 = n0;
            }

            tc::mbar_wait(&tfull[acc], acc_phase);
            __syncwarp();                 // reconverge before the .sync.aligned TMEM loads
            tc::tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
            epilogue_drain<EPI>(t_addr, stg, s_bias, ep, &tmC, n0, row0, ba, half * (BN / 2), (half + 1) * (BN / 2), lane);
            // every TMEM load of this warp has completed (last tmem_ld_wait inside): hand the accumulator back
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
        }
        if (lane == 0) tc::tma_store_wait<0>();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.  Each CTA stages its
// own 128 rows of A and HALF of the W tile (BN/2 rows); one MMA issued by the pair's leader reads both CTAs' shared memory
// and writes each CTA's 128 accumulator rows into its own TMEM.  Per CTA and k-block this moves 32 KB instead of 48 KB
// through L2 / the shared-memory ports for the same number of MACs, which is what lifts the 1-CTA kernel's ~66 % ceiling
// (UMMA operand reads + TMA fills exceed 128 B/clk/SM there).
// MODE 0: plain epilogues, 6-stage operand ring.
// MODE 1 (RES): the residual-adding kernels give two ring stages (64 KB) to a per-warp landing zone for the residual tile (cp.async).
// MODE 2 (FOLD): the consuming kernels give one stage (32 KB) to bias / colsum tables of ALL N columns (no per-tile refill, no
//         barrier per tile, no global loads in the drain) and to a per-warp landing zone for the row statistics of the next tile.
constexpr int FOLD_MAX_N = 1536;
template <int BN, int MODE = 0> struct Gemm2Cfg {
#ifndef SNB_RES_STAGES
#define SNB_RES_STAGES 4
#endif
    static constexpr int STAGES = MODE == 1 ? SNB_RES_STAGES : (MODE == 2 ? 5 : 6);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
    static constexpr int RESID_WARP_BYTES = 2 * STAGING_BYTES;                    // two 64-column chunks of 32 rows
    static constexpr int OFF_RESID = OFF_STAGING + EPI_WARPS * STAGING_BYTES;
    static constexpr int OFF_COLT = OFF_RESID + (MODE == 1 ? EPI_WARPS * RESID_WARP_BYTES : 0);     // MODE 2: bias[FOLD_MAX_N], colsum[FOLD_MAX_N]
    static constexpr int STATS_WARP_BYTES = 2 * 32 * 32;                          // two tiles x 32 rows x (4 x float2)
    static constexpr int OFF_STATS = OFF_COLT + (MODE == 2 ? 2 * FOLD_MAX_N * 4 : 0);
    static constexpr int OFF_BIAS = OFF_STATS + (MODE == 2 ? EPI_WARPS * STATS_WARP_BYTES : 0);
    static constexpr int OFF_BARS = OFF_BIAS + BN * 4;
    static constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024;
};
static_assert(Gemm2Cfg<256, 2>::SMEM_BYTES <= 227 * 1024 && Gemm2Cfg<256, 1>::SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Drain of one warp's 32 rows x 128 columns for the residual-adding kernels.  The residual tile is NOT loaded by this function: it
// arrives by cp.async in the warp's landing zone `rbuf` (same swizzled "row r, 16-byte piece c" layout as the staging tile, so that
// "thread = row" reads are conflict free), one commit group per 64-column chunk, issued a whole chunk-drain ahead: as soon as chunk
// ci of THIS tile has been read out, `next_chunk(ci)` issues chunk ci of the warp's NEXT tile into the same bytes.  The invariant
// "exactly one younger group in flight" makes every wait a cp.async.wait_group 1.
// Before: the 64 KB residual of a tile was fetched into registers at the top of the tile's epilogue and needed a few hundred clocks
// later; with the accumulator already waiting (the RES kernels are not MMA bound) its whole DRAM latency sat on the epilogue's
// critical path every tile -- ncu: DRAM 49 %, tensor pipe 34 %, neither bound.
template <int EPI, bool FOLD, int RES, class NextChunk>
__device__ __forceinline__ void epilogue_drain_res(uint32_t t_addr, uint8_t *stg, const uint8_t *rbuf, const float *s_bias, const GemmEpi &ep,
                                                   const CUtensorMap *tmC, int n0, int row0, int ba, int c_begin, int lane, RowFold &rf,
                                                   NextChunk &&next_chunk)
{
    uint32_t ra[32], rb[32], pk[16];
    uint64_t v[16];
    tc::tmem_ld_32x32(t_addr + c_begin, ra);
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
        const int c = c_begin + ci * 64;
        uint4 mine[8];
        cp_async_wait<1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) mine[j] = *staging_slot(const_cast<uint8_t *>(rbuf) + ci * STAGING_BYTES, lane, j);
        __syncwarp();                                   // every lane has its row: the chunk's bytes may be refilled
        next_chunk(ci);
        tc::tmem_ld_wait();
        tc::tmem_ld_32x32(t_addr + c + 32, rb);
        epilogue_math2<EPI, FOLD>(ra, v, s_bias + c, ep, n0 + c, ba, rf);
        {
            const uint4 r4[4] = {mine[0], mine[1], mine[2], mine[3]};
            residual_pack2<RES>(v, r4, ep, n0 + c, rf, pk);
        }
        if (lane == 0) tc::tma_store_wait_read<0>();    // the previous TMA store has read the output staging tile
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) *staging_slot(stg, lane, j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tc::tmem_ld_wait();
        if (ci == 0) tc::tmem_ld_32x32(t_addr + c + 64, ra);
        epilogue_math2<EPI, FOLD>(rb, v, s_bias + c + 32, ep, n0 + c + 32, ba, rf);
        {
            const uint4 r4[4] = {mine[4], mine[5], mine[6], mine[7]};
            residual_pack2<RES>(v, r4, ep, n0 + c + 32, rf, pk);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *staging_slot(stg, lane, 4 + j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tc::tma_store_2d(tmC, stg, n0 + c, row0); tc::tma_store_commit(); }
    }
}

//
// LayerNorm fold (FOLD / RES template parameters; BN = 256 only).  The post-norm encoder layer is  y = LN1(h + attn(h)),
// h' = LN2(y + ff(y)).  Instead of a LayerNorm kernel between the GEMMs (two HBM passes per layer, 11.7 % of the r01 step):
//   * a PRODUCING GEMM (out-proj, linear2; RES != 0) adds the residual in its epilogue, writes the PRE-norm row z (bf16) and the
//     row's sum / sum of squares (of the rounded values, one float2 per (row, 128-column slice));  RES == 2: the residual itself is a
//     LayerNorm output that was never materialised -- it is recomputed from the previous z and ITS statistics, gamma, beta;
//   * a CONSUMING GEMM (in_proj, linear1, concat3; FOLD) multiplies z by W' = W . diag(gamma) (folded once at model load) and undoes
//     the normalisation per row in its epilogue:  LN(z) W^T = rstd (z W'^T) - rstd mu colsum(W') + (beta W^T + b).
// Each epilogue thread owns one accumulator row, so mu / rstd are per-thread scalars; the residual slice of the thread (256 B) is
// prefetched into registers before the accumulator is waited for.
template <int BN, int EPI, bool FOLD = false, int RES = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const GemmEpi ep, const int M, const int N, const int K)
{
    using Cfg = Gemm2Cfg<BN, RES != 0 ? 1 : (FOLD ? 2 : 0)>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::OFF_BARS);
    uint64_t *full = bars, *empty = bars + Cfg::STAGES, *tfull = bars + 2 * Cfg::STAGES, *tempty = bars + 2 * Cfg::STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * Cfg::STAGES + 4);
    float *s_bias = reinterpret_cast<float *>(smem + Cfg::OFF_BIAS);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;
    const int m_tiles = (M + 2 * BM - 1) / (2 * BM), n_tiles = N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = K / BK;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA); tc::prefetch_tmap(&tmB); tc::prefetch_tmap(&tmC);
        for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }   // full: armed by the leader for both CTAs' bytes
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 2 * EPI_WARPS); }  // tempty: both epilogues
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();          // the peer's barriers are initialised before anything arrives on them
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one per CTA) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += n_pairs) {
                const int m0 = (t / n_tiles) * (2 * BM) + (int)rank * BM, n0 = (t % n_tiles) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *a = smem + stage * Cfg::STAGE_BYTES;
                    // the leader arms its barrier with the bytes of BOTH CTAs; the peer's loads complete_tx on the same barrier
                    // (its stage cannot be refilled before the leader's previous phase finished: empty[] is signalled by
                    // the commit that follows the MMAs which waited on that phase)
                    if (leader) tc::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                    tc::tma_load_2d_pair(a, &tmA, &full[stage], kb * BK, m0);
                    tc::tma_load_2d_pair(a + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        // (issue cost: see the single-CTA kernel -- uniform warp index / TMEM base, one elected thread, precomputed descriptors)
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        if (leader && tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(2 * BM, BN, 0, 0);
            const uint64_t da0 = tc::make_smem_desc_sw128(tc::smem_u32(smem), 16, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = pair; t < num_tiles; t += n_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_u + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint64_t da = da0 + (uint64_t)stage * (Cfg::STAGE_BYTES >> 4), db = da + (Cfg::A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc::umma_ss_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc::umma_commit_pair(&empty[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit_pair(&tfull[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 of both CTAs; each CTA drains its own 128 accumulator rows) ==========
        const int quarter = warp & 3;            // TMEM lanes [32*quarter, 32*quarter + 32)
        const int half = (warp - 2) >> 2;        // which half of the tile's BN columns this warp drains
        const int etid = (warp - 2) * 32 + lane;
        uint8_t *stg = smem + Cfg::OFF_STAGING + (warp - 2) * STAGING_BYTES;
        uint8_t *rbuf = smem + Cfg::OFF_RESID + (warp - 2) * Cfg::RESID_WARP_BYTES;      // RES != 0 only
        // chunk (0 / 1 = the two 64-column halves of this warp's 128 columns) of the residual of tile t -> landing zone, coalesced:
        // per instruction the warp copies 4 rows x 128 B (lane = 16-byte piece (lane & 7) of row 4 i + (lane >> 3))
        auto issue_resid = [&](int t, int chunk) {
            const int tm0 = (t / n_tiles) * (2 * BM) + (int)rank * BM + quarter * 32, tn0 = (t % n_tiles) * BN + half * (BN / 2) + chunk * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int rr = tm0 + i * 4 + (lane >> 3);
                if (rr >= M) rr = M - 1;
                cp_async_16(staging_slot(rbuf + chunk * STAGING_BYTES, i * 4 + (lane >> 3), lane & 7),
                            reinterpret_cast<const uint4 *>(ep.resid + (size_t)rr * N + tn0) + (lane & 7));
            }
            cp_async_commit();
        };
        auto load_res_stats = [&](int t, float4 &a, float4 &b) {
            int row = (t / n_tiles) * (2 * BM) + (int)rank * BM + quarter * 32 + lane;
            if (row >= M) row = M - 1;
            const float4 *sp = reinterpret_cast<const float4 *>(ep.res_stats + (size_t)row * 4);
            a = __ldg(sp); b = __ldg(sp + 1);
        };
        // FOLD: bias / colsum of all N columns -> shared memory once; row statistics of the NEXT tile by cp.async (32 bytes per row, each
        // thread lands and later reads its own row: no warp synchronisation needed, one commit group per tile, wait_group 1)
        float *s_ball = reinterpret_cast<float *>(smem + Cfg::OFF_COLT), *s_call = s_ball + FOLD_MAX_N;
        uint8_t *sbuf = smem + Cfg::OFF_STATS + (warp - 2) * Cfg::STATS_WARP_BYTES;
        auto issue_stats = [&](int t, int buf) {
            int row = (t / n_tiles) * (2 * BM) + (int)rank * BM + quarter * 32 + lane;
            if (row >= M) row = M - 1;
            const float4 *sp = reinterpret_cast<const float4 *>(ep.stats_in + (size_t)row * 4);
            cp_async_16(sbuf + buf * 1024 + lane * 32, sp);
            cp_async_16(sbuf + buf * 1024 + lane * 32 + 16, sp + 1);
            cp_async_commit();
        };
        if constexpr (FOLD) {
            for (int i = etid; i < N; i += EPI_WARPS * 32) { s_ball[i] = __ldg(ep.bias + i); s_call[i] = __ldg(ep.colsum + i); }
            tc::named_bar_sync(1, EPI_WARPS * 32);
            if (pair < num_tiles) issue_stats(pair, 0); else cp_async_commit();
        }
        float4 nsa = make_float4(0.f, 0.f, 0.f, 0.f), nsb = nsa;          // RES == 2: row statistics of the residual, one tile ahead
        if constexpr (RES != 0) {
            if (pair < num_tiles) {
                issue_resid(pair, 0); issue_resid(pair, 1);
                if constexpr (RES == 2) load_res_stats(pair, nsa, nsb);
            } else { cp_async_commit(); cp_async_commit(); }
        }
        int it = 0, bias_n0 = -1;
        for (int t = pair; t < num_tiles; t += n_pairs, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (t / n_tiles) * (2 * BM) + (int)rank * BM, n0 = (t % n_tiles) * BN;
            const int row0 = m0 + quarter * 32;
            int ba = 0;
            if constexpr (EPI == EPI_CSL_BF16) {
                int row = row0 + lane;
                if (row >= M) row = M - 1;
                const int b = row / ep.tok_per_env;
                const int r = (row - b * ep.tok_per_env) / ep.T;
                ba = b * ep.A + (r % ep.A);
            }
            // this tile's bias slice -> smem, only when the column block changed (a pair of a 2-column-block GEMM keeps its block for
            // the whole launch: n_pairs is even); the previous tile's readers are past their last use: barrier below
            if (!FOLD && n0 != bias_n0) {
                tc::named_bar_sync(1, EPI_WARPS * 32);
                for (int i = etid; i < BN; i += EPI_WARPS * 32) {
                    float b = __ldg(ep.bias + n0 + i);
                    if constexpr (RES == 2) b += __ldg(ep.res_beta + n0 + i);      // beta of the recomputed LayerNorm residual rides with the bias
                    s_bias[i] = b;
                }
                tc::named_bar_sync(1, EPI_WARPS * 32);
                bias_n0 = n0;
            }

            // LayerNorm fold: this thread's row statistics, loaded before the accumulator wait
            RowFold rf = RowFold();
            rf.sum2 = tc::f2_pack(0.0f, 0.0f); rf.sq2 = rf.sum2;
            const int t_next = t + n_pairs;
            const bool has_next = t_next < num_tiles;
            {
                if constexpr (FOLD) {
                    if (has_next) issue_stats(t_next, (it + 1) & 1); else cp_async_commit();
                    cp_async_wait<1>();                                       // this tile's statistics (requested a tile ago) have landed
                    const float4 a = *reinterpret_cast<const float4 *>(sbuf + (it & 1) * 1024 + lane * 32);
                    const float4 b = *reinterpret_cast<const float4 *>(sbuf + (it & 1) * 1024 + lane * 32 + 16);
                    const float mu = (a.x + a.z + b.x + b.z) * (1.0f / 512.0f);
                    const float var = (a.y + a.w + b.y + b.w) * (1.0f / 512.0f) - mu * mu;
                    rf.rstd = rsqrtf(fmaxf(var, 0.0f) + 1e-5f);
                    rf.nmr = -rf.rstd * mu;
                }
                if constexpr (RES == 2) {
                    rf.r_mu = (nsa.x + nsa.z + nsb.x + nsb.z) * (1.0f / 512.0f);
                    const float var = (nsa.y + nsa.w + nsb.y + nsb.w) * (1.0f / 512.0f) - rf.r_mu * rf.r_mu;
                    rf.r_rstd = rsqrtf(fmaxf(var, 0.0f) + 1e-5f);
                    if (has_next) load_res_stats(t_next, nsa, nsb);           // in flight during this tile's drain
                }
            }
            tc::mbar_wait(&tfull[acc], acc_phase);
            __syncwarp();                 // reconverge before the .sync.aligned TMEM loads
            tc::tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
            if constexpr (RES != 0) {
                static_assert(BN == 256, "the LayerNorm fold is written for 256-column tiles (two 64-column chunks per warp)");
                epilogue_drain_res<EPI, FOLD, RES>(t_addr, stg, rbuf, s_bias, ep, &tmC, n0, row0, ba, half * (BN / 2), lane, rf,
                                                   [&](int chunk) { if (has_next) issue_resid(t_next, chunk); else cp_async_commit(); });
            } else if constexpr (FOLD) {
                static_assert(BN == 256, "the LayerNorm fold is written for 256-column tiles (two 64-column chunks per warp)");
                epilogue_drain_fold<EPI, FOLD>(t_addr, stg, s_ball + n0, s_call + n0, ep, &tmC, n0, row0, ba, half * (BN / 2), lane, rf);
            } else {
                epilogue_drain<EPI>(t_addr, stg, s_bias, ep, &tmC, n0, row0, ba, half * (BN / 2), (half + 1) * (BN / 2), lane);
            }
            // every TMEM load of this warp has completed (last tmem_ld_wait inside): hand the accumulator back
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) tc::mbar_arrive(&tempty[acc]);
                else tc::mbar_arrive_cluster(tc::mapa_u32(&tempty[acc], 0));
            }
            if constexpr (RES != 0) {
                float s_lo, s_hi, q_lo, q_hi;
                tc::f2_unpack(rf.sum2, s_lo, s_hi); tc::f2_unpack(rf.sq2, q_lo, q_hi);
                if (row0 + lane < M) ep.stats_out[(size_t)(row0 + lane) * 4 + (n0 / BN) * 2 + half] = make_float2(s_lo + s_hi, q_lo + q_hi);
            }
        }
        if constexpr (RES != 0 || FOLD) cp_async_wait<0>();
        if (lane == 0) tc::tma_store_wait<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();          // nobody leaves while the pair's MMAs / commits may still touch its shared memory
    if (warp == 1) tc::tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

template <int BN, int EPI>
int launch_t(const GemmPlan *p, const GemmEpi *ep, int num_sms, cudaStream_t stream)
{
    using Cfg = GemmCfg<BN>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    });
    SNB_CUDA_TRY(attr_err);
    SNB_REQUIRE((EPI == EPI_BIAS_F32) == (p->out_f32 != 0), SNB_EINVAL, "gemm: epilogue %d does not match the plan's output type", EPI);
    const int tiles = ((p->M + BM - 1) / BM) * (p->N / BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    gemm_bf16_tn_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(p->tmA, p->tmB, p->tmC, *ep, p->M, p->N, p->K);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

template <int BN, int EPI, bool FOLD = false, int RES = 0>
int launch2_t(const GemmPlan *p, const GemmEpi *ep, int num_sms, cudaStream_t stream)
{
    using Cfg = Gemm2Cfg<BN, RES != 0 ? 1 : (FOLD ? 2 : 0)>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BN, EPI, FOLD, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    });
    SNB_CUDA_TRY(attr_err);
    const int tiles = ((p->M + 2 * BM - 1) / (2 * BM)) * (p->N / BN);
    const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
    if (FOLD) SNB_REQUIRE(ep->colsum && ep->stats_in && p->K == 512 && p->N <= FOLD_MAX_N, SNB_EINVAL, "gemm: LayerNorm fold needs colsum / stats_in, K = 512 and N <= 1536");
    if (RES) SNB_REQUIRE(ep->resid && ep->stats_out && p->N == 512 && (RES == 1 || (ep->res_stats && ep->res_gamma && ep->res_beta)), SNB_EINVAL,
                         "gemm: residual epilogue needs resid / stats_out (N = 512) and, for a normalised residual, its stats / gamma / beta");
    gemm2_bf16_tn_kernel<BN, EPI, FOLD, RES><<<2 * pairs, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(p->tmA, p->tmB2, p->tmC, *ep, p->M, p->N, p->K);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

bool use_pair_kernel()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("SNB_GEMM_2CTA"); v = e ? atoi(e) : 1; }
    return v != 0;
}

template <int BN>
int launch_bn(const GemmPlan *p, int kind, const GemmEpi *ep, int num_sms, cudaStream_t stream)
{
    if (ep->fold || ep->res) {
        SNB_REQUIRE(BN == 256 && use_pair_kernel(), SNB_EUNSUPPORTED, "gemm: the LayerNorm-fold epilogues exist for the CTA-pair kernel (N %% 256 == 0) only");
        if (kind == EPI_BIAS_BF16 && ep->fold && !ep->res) return launch2_t<256, EPI_BIAS_BF16, true, 0>(p, ep, num_sms, stream);
        if (kind == EPI_BIAS_RELU_BF16 && ep->fold && !ep->res) return launch2_t<256, EPI_BIAS_RELU_BF16, true, 0>(p, ep, num_sms, stream);
        if (kind == EPI_CSL_BF16 && ep->fold && !ep->res) return launch2_t<256, EPI_CSL_BF16, true, 0>(p, ep, num_sms, stream);
        if (kind == EPI_BIAS_BF16 && !ep->fold && ep->res == 1) return launch2_t<256, EPI_BIAS_BF16, false, 1>(p, ep, num_sms, stream);
        if (kind == EPI_BIAS_BF16 && !ep->fold && ep->res == 2) return launch2_t<256, EPI_BIAS_BF16, false, 2>(p, ep, num_sms, stream);
        snb_set_error("gemm: unsupported LayerNorm-fold combination (epilogue %d, fold %d, res %d)", kind, ep->fold, ep->res);
        return SNB_EUNSUPPORTED;
    }
    if (BN == 256 && use_pair_kernel() && kind != EPI_BIAS_F32) {
        switch (kind) {
        case EPI_BIAS_BF16: return launch2_t<256, EPI_BIAS_BF16>(p, ep, num_sms, stream);
        case EPI_BIAS_RELU_BF16: return launch2_t<256, EPI_BIAS_RELU_BF16>(p, ep, num_sms, stream);
        case EPI_CSL_BF16: return launch2_t<256, EPI_CSL_BF16>(p, ep, num_sms, stream);
        }
    }
    switch (kind) {
    case EPI_BIAS_BF16: return launch_t<BN, EPI_BIAS_BF16>(p, ep, num_sms, stream);
    case EPI_BIAS_RELU_BF16: return launch_t<BN, EPI_BIAS_RELU_BF16>(p, ep, num_sms, stream);
    case EPI_BIAS_F32: return launch_t<BN, EPI_BIAS_F32>(p, ep, num_sms, stream);
    case EPI_CSL_BF16: return launch_t<BN, EPI_CSL_BF16>(p, ep, num_sms, stream);
    }
    snb_set_error("gemm: unknown epilogue %d", kind);
    return SNB_EINVAL;
}

int make_tmap(CUtensorMap *out, CUtensorMapDataType dt, int elem_bytes, int rank, const void *ptr, const cuuint64_t *dims,
              const cuuint32_t *box)
{
    EncodeTiledFn enc = get_encode();
    SNB_REQUIRE(enc != nullptr, SNB_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    SNB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (dims[0] * elem_bytes) % 16 == 0, SNB_EINVAL, "tensor map: unaligned tensor");
    cuuint64_t strides[2] = {dims[0] * elem_bytes, dims[0] * dims[1] * elem_bytes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, dt, rank, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SNB_REQUIRE(r == CUDA_SUCCESS, SNB_ECUDA, "cuTensorMapEncodeTiled(rank %d) failed: %d", rank, (int)r);
    return SNB_OK;
}

} // namespace

int snb_make_tmap_2d(CUtensorMap *out, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint32_t box[2] = {64, box_rows};
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2, ptr, dims, box);
}

int snb_make_tmap_3d(CUtensorMap *out, const void *ptr, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box_rows)
{
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint32_t box[3] = {64, box_rows, 1};
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, ptr, dims, box);
}

int snb_gemm_plan(GemmPlan *plan, const bf16 *A, const bf16 *W, void *out, int out_f32, int M, int N, int K)
{
    SNB_REQUIRE(M > 0 && K % BK == 0 && N % 128 == 0, SNB_EUNSUPPORTED, "gemm: unsupported shape M=%d N=%d K=%d", M, N, K);
    plan->M = M; plan->N = N; plan->K = K; plan->out_f32 = out_f32;
    plan->BN = (N % 256 == 0) ? 256 : 128;
    int rc = snb_make_tmap_2d(&plan->tmA, A, (uint64_t)M, (uint64_t)K, BM);
    if (rc) return rc;
    rc = snb_make_tmap_2d(&plan->tmB, W, (uint64_t)N, (uint64_t)K, (uint32_t)plan->BN);
    if (rc) return rc;
    rc = snb_make_tmap_2d(&plan->tmB2, W, (uint64_t)N, (uint64_t)K, (uint32_t)plan->BN / 2);   // CTA-pair kernel: half of W per CTA
    if (rc) return rc;
    // output [M, N]: store unit = 32 rows x 128 bytes
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint32_t box[2] = {out_f32 ? 32u : 64u, 32u};
    return make_tmap(&plan->tmC, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out_f32 ? 4 : 2, 2, out, dims, box);
}

int snb_gemm_launch(const GemmPlan *plan, int epi_kind, const GemmEpi *epi, int num_sms, cudaStream_t stream)
{
    return plan->BN == 256 ? launch_bn<256>(plan, epi_kind, epi, num_sms, stream) : launch_bn<128>(plan, epi_kind, epi, num_sms, stream);
}
