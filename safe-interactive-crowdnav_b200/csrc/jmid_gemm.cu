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
#include <mutex>

#include "jmid_internal.h"
#include "tc_utils.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int STAGING_BYTES = 32 * 128; // per epilogue warp: 32 rows x 128 B (one swizzle atom wide)

template <int BN> struct GemmCfg {
    static constexpr int STAGES = 4;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN; // 512 or 256: both powers of two
    static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_STAGING + 4 * STAGING_BYTES;
    static constexpr int OFF_BARS = OFF_BIAS + BN * 4;
    static constexpr int SMEM_BYTES = OFF_BARS + 256 + 1024 /*align*/;
};

// one 32-column chunk of one accumulator row: + bias (+ReLU | ConcatSquash), result left in v[]
template <int EPI>
__device__ __forceinline__ void epilogue_math(const uint32_t (&acc)[32], float (&v)[32], const float *s_bias, const GemmEpi &ep,
                                              int col0_global, int ba)
{
    const float4 *b4 = reinterpret_cast<const float4 *>(s_bias);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = b4[j];
        v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
        v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
        v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
        v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
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

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmEpi ep, const int M, const int N, const int K)
{
    using Cfg = GemmCfg<BN>;
    constexpr bool OUT_F32 = (EPI == EPI_BIAS_F32);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::OFF_BARS);
    uint64_t *full = bars, *empty = bars + Cfg::STAGES, *tfull = bars + 2 * Cfg::STAGES, *tempty = bars + 2 * Cfg::STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * Cfg::STAGES + 4);
    float *s_bias = reinterpret_cast<float *>(smem + Cfg::OFF_BIAS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + BM - 1) / BM, n_tiles = N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int k_blocks = K / BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        tc::prefetch_tmap(&tmC);
        for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 128); }
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
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t a_addr = tc::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = tc::make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t db = tc::make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        tc::umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc::umma_commit(&empty[stage]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3; // TMEM lanes [32*quarter, 32*quarter + 32)
        const int etid = (warp - 2) * 32 + lane;
        uint8_t *stg = smem + Cfg::OFF_STAGING + quarter * STAGING_BYTES;
        int it = 0;
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
            // this tile's bias slice -> smem (the previous tile's readers are past their last use: barrier below)
            tc::named_bar_sync(1, 128);
            for (int i = etid; i < BN; i += 128) s_bias[i] = __ldg(ep.bias + n0 + i);
            tc::named_bar_sync(1, 128);

            tc::mbar_wait(&tfull[acc], acc_phase);
            __syncwarp();                 // reconverge before the .sync.aligned TMEM loads
            tc::tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
            uint32_t ra[32], rb[32];
            float v[32];
            tc::tmem_ld_32x32(t_addr, ra);
            if constexpr (OUT_F32) {
                // fp32 output: one 32-column chunk = 128 B per row = one store unit
#pragma unroll 1
                for (int c = 0; c < BN / 32; c += 2) {
                    tc::tmem_ld_wait();
                    tc::tmem_ld_32x32(t_addr + (c + 1) * 32, rb);
                    epilogue_math<EPI>(ra, v, s_bias + c * 32, ep, n0 + c * 32, ba);
                    if (lane == 0) tc::tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) *staging_slot(stg, lane, j) = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tc::tma_store_2d(&tmC, stg, n0 + c * 32, row0); tc::tma_store_commit(); }
                    tc::tmem_ld_wait();
                    if (c + 2 < BN / 32) tc::tmem_ld_32x32(t_addr + (c + 2) * 32, ra);
                    epilogue_math<EPI>(rb, v, s_bias + (c + 1) * 32, ep, n0 + (c + 1) * 32, ba);
                    if (lane == 0) tc::tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) *staging_slot(stg, lane, j) = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tc::tma_store_2d(&tmC, stg, n0 + (c + 1) * 32, row0); tc::tma_store_commit(); }
                }
            } else {
                // bf16 output: two 32-column chunks = 128 B per row = one store unit
#pragma unroll 1
                for (int c = 0; c < BN / 32; c += 2) {
                    tc::tmem_ld_wait();
                    tc::tmem_ld_32x32(t_addr + (c + 1) * 32, rb);
                    epilogue_math<EPI>(ra, v, s_bias + c * 32, ep, n0 + c * 32, ba);
                    if (lane == 0) tc::tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *staging_slot(stg, lane, j) = make_uint4(tc::pack_bf16(v[8 * j], v[8 * j + 1]), tc::pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                                 tc::pack_bf16(v[8 * j + 4], v[8 * j + 5]), tc::pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    tc::tmem_ld_wait();
                    if (c + 2 < BN / 32) tc::tmem_ld_32x32(t_addr + (c + 2) * 32, ra);
                    epilogue_math<EPI>(rb, v, s_bias + (c + 1) * 32, ep, n0 + (c + 1) * 32, ba);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *staging_slot(stg, lane, 4 + j) = make_uint4(tc::pack_bf16(v[8 * j], v[8 * j + 1]), tc::pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                                     tc::pack_bf16(v[8 * j + 4], v[8 * j + 5]), tc::pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tc::tma_store_2d(&tmC, stg, n0 + c * 32, row0); tc::tma_store_commit(); }
                }
            }
            // every TMEM load of this tile has completed (last tmem_ld_wait above): hand the accumulator back
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty[acc]);
        }
        if (lane == 0) tc::tma_store_wait<0>();
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
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

template <int BN>
int launch_bn(const GemmPlan *p, int kind, const GemmEpi *ep, int num_sms, cudaStream_t stream)
{
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
    // output [M, N]: store unit = 32 rows x 128 bytes
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint32_t box[2] = {out_f32 ? 32u : 64u, 32u};
    return make_tmap(&plan->tmC, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out_f32 ? 4 : 2, 2, out, dims, box);
}

int snb_gemm_launch(const GemmPlan *plan, int epi_kind, const GemmEpi *epi, int num_sms, cudaStream_t stream)
{
    return plan->BN == 256 ? launch_bn<256>(plan, epi_kind, epi, num_sms, stream) : launch_bn<128>(plan, epi_kind, epi, num_sms, stream);
}
