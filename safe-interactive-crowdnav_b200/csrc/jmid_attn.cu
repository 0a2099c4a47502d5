// csrc/jmid_attn.cu -- fused multi-head self-attention of the JMID noise network on tcgen05.
//
// The reference flattens the T*A*S tokens of one environment into ONE unmasked sequence and runs
// nn.MultiheadAttention(d=512, heads=4) over it (models/diffusion.py:196-204, quirk q1).  This kernel computes, per
// (environment, head, 128-query tile):  O = softmax(Q K^T / sqrt(128)) V  flash-style, never materialising the
// N x N score matrix in HBM:
//   warp 0      TMA producer: Q tile once, then K / V blocks of 128 keys through 2-stage smem rings
//   warp 1      MMA issuer:   S = Q K^T (SS, M=128 N<=128 K=128) into a double-buffered TMEM S tile,
//                             O += P V (TS: P read from TMEM as bf16, V MN-major from smem) into a TMEM O tile
//   warp 2      TMEM allocator
//   warps 4..7  softmax:      one thread per query row: tcgen05.ld the S row, running max with lazy rescale of O,
//                             exp2, bf16 P written back over S with tcgen05.st, final O / l -> global
// S(j+1) is issued before waiting for P(j), so the QK^T of the next block overlaps the softmax of the current one.
#include <cfloat>
#include <cstdlib>
#include <mutex>

#include "jmid_internal.h"
#include "tc_utils.cuh"

namespace {

constexpr int HD = 128;       // head dim
constexpr int NHEAD = 4;
constexpr int BQ = 128;       // queries per CTA
constexpr int BKV = 128;      // keys per block
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = BQ * HD * 2; // 32 KB: two 64-column boxes of 16 KB
constexpr int ATTN_THREADS = 256;
constexpr int ATTN_SMEM = TILE_BYTES * (1 + 2 * KV_STAGES) + 1024 + 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O = 256;
constexpr float RESCALE_THRESHOLD = 8.0f; // log2 units: P <= 2^8 before the running max is refreshed

struct AttnArgs {
    bf16 *out;        // [n_env * n_tok, 512]
    int n_tok;
    float scale_log2; // log2(e) / sqrt(HD)
};

__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;
    uint8_t *sK = smem + TILE_BYTES;
    uint8_t *sV = smem + TILE_BYTES * (1 + KV_STAGES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TILE_BYTES * (1 + 2 * KV_STAGES));
    uint64_t *q_full = bars;
    uint64_t *k_full = bars + 1, *k_empty = bars + 3, *v_full = bars + 5, *v_empty = bars + 7;
    uint64_t *s_full = bars + 9, *p_ready = bars + 11, *pv_done = bars + 13;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, head = blockIdx.y, env = blockIdx.z;
    const int n_tok = args.n_tok;
    const int n_kv = (n_tok + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQKV);
        tc::mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1);
            tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
            tc::mbar_init(&s_full[s], 1); tc::mbar_init(&p_ready[s], 128); tc::mbar_init(&pv_done[s], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int cq = head * HD, ck = 512 + head * HD, cv = 1024 + head * HD;
            tc::mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tc::tma_load_3d(sQ, &tmQKV, q_full, cq, q0, env);
            tc::tma_load_3d(sQ + TILE_BYTES / 2, &tmQKV, q_full, cq + 64, q0, env);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                tc::mbar_wait(&k_empty[st], ph ^ 1);
                tc::mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
                tc::tma_load_3d(sK + st * TILE_BYTES, &tmQKV, &k_full[st], ck, j * BKV, env);
                tc::tma_load_3d(sK + st * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &k_full[st], ck + 64, j * BKV, env);
                tc::mbar_wait(&v_empty[st], ph ^ 1);
                tc::mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
                tc::tma_load_3d(sV + st * TILE_BYTES, &tmQKV, &v_full[st], cv, j * BKV, env);
                tc::tma_load_3d(sV + st * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &v_full[st], cv + 64, j * BKV, env);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t q_addr = tc::smem_u32(sQ);
            auto kv_cols = [&](int j) { // keys of block j rounded up to the UMMA N granularity (16)
                const int rem = n_tok - j * BKV;
                return rem >= BKV ? BKV : ((rem + 15) & ~15);
            };
            auto issue_S = [&](int j) {
                const int st = j & 1;
                tc::mbar_wait(&k_full[st], (j >> 1) & 1);
                tc::tc_fence_after();
                const uint32_t k_addr = tc::smem_u32(sK + st * TILE_BYTES);
                const uint32_t idesc = tc::make_idesc_bf16(BQ, (uint32_t)kv_cols(j), 0, 0);
                const uint32_t d = tmem_base + (st ? TM_S1 : TM_S0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint32_t off = (k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32;
                    tc::umma_ss(d, tc::make_smem_desc_sw128(q_addr + off, 16, 1024), tc::make_smem_desc_sw128(k_addr + off, 16, 1024),
                                idesc, k != 0 ? 1u : 0u);
                }
                tc::umma_commit(&k_empty[st]);
                tc::umma_commit(&s_full[st]);
            };
            tc::mbar_wait(q_full, 0);
            tc::tc_fence_after();
            issue_S(0);
            constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, HD, 0, 1); // B = V is MN-major (head dim contiguous)
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                if (j + 1 < n_kv) issue_S(j + 1);
                tc::mbar_wait(&p_ready[st], ph);
                tc::mbar_wait(&v_full[st], ph);
                tc::tc_fence_after();
                const uint32_t v_addr = tc::smem_u32(sV + st * TILE_BYTES);
                const uint32_t p_tmem = tmem_base + (st ? TM_S1 : TM_S0);
                const int ksteps = kv_cols(j) / 16;
                for (int k = 0; k < ksteps; ++k) {
                    // 16 keys = 2 groups of 8 rows (SBO 1024 B); the two 64-wide head-dim boxes are LBO = 16 KB apart
                    const uint64_t dv = tc::make_smem_desc_sw128(v_addr + k * 2048, TILE_BYTES / 2, 1024);
                    tc::umma_ts(tmem_base + TM_O, p_tmem + k * 8, dv, idesc_pv, (j | k) != 0 ? 1u : 0u);
                }
                tc::umma_commit(&v_empty[st]);
                tc::umma_commit(&pv_done[st]);
            }
        }
    } else if (warp >= 4) {
        // ===================== softmax / correction / epilogue =====================
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const float c = args.scale_log2;
        float m_used = -INFINITY; // running max the exponents are taken against (raw score units)
        float l = 0.0f;
        for (int j = 0; j < n_kv; ++j) {
            const int st = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            tc::mbar_wait(&s_full[st], ph);
            __syncwarp();                 // reconverge before the .sync.aligned TMEM loads
            tc::tc_fence_after();
            const uint32_t s_addr = tmem_base + lane_addr + (st ? TM_S1 : TM_S0);
            // the whole S row: four 32-column loads in flight, ONE tcgen05.wait::ld
            uint32_t s0[32], s1[32], s2[32], s3[32];
            tc::tmem_ld_32x32(s_addr, s0);
            tc::tmem_ld_32x32(s_addr + 32, s1);
            tc::tmem_ld_32x32(s_addr + 64, s2);
            tc::tmem_ld_32x32(s_addr + 96, s3);
            tc::tmem_ld_wait();
            const int valid = n_tok - j * BKV; // keys of this block that exist
            if (valid < BKV) {                 // ragged last block: keys that do not exist score -inf
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= valid) s0[i] = 0xff800000u;
                    if (32 + i >= valid) s1[i] = 0xff800000u;
                    if (64 + i >= valid) s2[i] = 0xff800000u;
                    if (96 + i >= valid) s3[i] = 0xff800000u;
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                mx0 = fmaxf(mx0, __uint_as_float(s0[i])); mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
                mx2 = fmaxf(mx2, __uint_as_float(s2[i])); mx3 = fmaxf(mx3, __uint_as_float(s3[i]));
            }
            const float bmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            if (j == 0) {
                m_used = bmax;
            } else if (__any_sync(0xffffffffu, (bmax - m_used) * c > RESCALE_THRESHOLD)) {
                // Refresh the running max.  The decision is WARP-uniform (tcgen05.ld/st are .sync.aligned and must be
                // executed by all 32 lanes together); lanes whose own max did not grow rescale by exactly 1.
                const int pst = (j - 1) & 1;
                tc::mbar_wait(&pv_done[pst], ((j - 1) >> 1) & 1);     // O is stable once the previous PV has landed
                __syncwarp();
                tc::tc_fence_after();
                const float m_new = fmaxf(m_used, bmax);
                const float f = tc::ex2_approx((m_used - m_new) * c);
                const uint32_t o_addr = tmem_base + lane_addr + TM_O;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t r[32];
                    tc::tmem_ld_32x32(o_addr + ch * 32, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                    tc::tmem_st_32x32(o_addr + ch * 32, r);
                }
                l *= f;
                m_used = m_new;
            }
            const float mc = m_used * c;
            float sum0 = 0.0f, sum1 = 0.0f, sum2 = 0.0f, sum3 = 0.0f;
            {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a0 = tc::ex2_approx(fmaf(__uint_as_float(s0[2 * i]), c, -mc)), a1 = tc::ex2_approx(fmaf(__uint_as_float(s0[2 * i + 1]), c, -mc));
                    const float b0 = tc::ex2_approx(fmaf(__uint_as_float(s1[2 * i]), c, -mc)), b1 = tc::ex2_approx(fmaf(__uint_as_float(s1[2 * i + 1]), c, -mc));
                    sum0 += a0; sum1 += a1; sum2 += b0; sum3 += b1;
                    pk[i] = tc::pack_bf16(a0, a1); pk[16 + i] = tc::pack_bf16(b0, b1);
                }
                tc::tmem_st_32x32(s_addr, pk);          // P columns [0,32)  <- S columns [0,64)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a0 = tc::ex2_approx(fmaf(__uint_as_float(s2[2 * i]), c, -mc)), a1 = tc::ex2_approx(fmaf(__uint_as_float(s2[2 * i + 1]), c, -mc));
                    const float b0 = tc::ex2_approx(fmaf(__uint_as_float(s3[2 * i]), c, -mc)), b1 = tc::ex2_approx(fmaf(__uint_as_float(s3[2 * i + 1]), c, -mc));
                    sum0 += a0; sum1 += a1; sum2 += b0; sum3 += b1;
                    pk[i] = tc::pack_bf16(a0, a1); pk[16 + i] = tc::pack_bf16(b0, b1);
                }
                tc::tmem_st_32x32(s_addr + 32, pk);     // P columns [32,64) <- S columns [64,128)
            }
            l += (sum0 + sum1) + (sum2 + sum3);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&p_ready[st]);
        }
        // final: O / l -> global
        const int lst = (n_kv - 1) & 1;
        tc::mbar_wait(&pv_done[lst], ((n_kv - 1) >> 1) & 1);
        __syncwarp();
        tc::tc_fence_after();
        const float inv_l = 1.0f / l;
        const int row = q0 + row_in_tile;
        const uint32_t o_addr = tmem_base + lane_addr + TM_O;
        bf16 *dst = args.out + ((size_t)env * n_tok + row) * (NHEAD * HD) + head * HD;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t r[32];
            tc::tmem_ld_32x32(o_addr + ch * 32, r);
            tc::tmem_ld_wait();
            if (row < n_tok) {
                uint4 *o4 = reinterpret_cast<uint4 *>(dst + ch * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o;
                    o.x = tc::pack_bf16(__uint_as_float(r[8 * q + 0]) * inv_l, __uint_as_float(r[8 * q + 1]) * inv_l);
                    o.y = tc::pack_bf16(__uint_as_float(r[8 * q + 2]) * inv_l, __uint_as_float(r[8 * q + 3]) * inv_l);
                    o.z = tc::pack_bf16(__uint_as_float(r[8 * q + 4]) * inv_l, __uint_as_float(r[8 * q + 5]) * inv_l);
                    o.w = tc::pack_bf16(__uint_as_float(r[8 * q + 6]) * inv_l, __uint_as_float(r[8 * q + 7]) * inv_l);
                    o4[q] = o;
                }
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// attn_fwd2_kernel: TWO 128-query tiles (A, B) of one (environment, head) per CTA, ping-ponged through the tensor pipe:
//   warp 0        TMA producer: Q_A, Q_B once, then K / V blocks of 128 keys through 2-stage rings (each K / V block is
//                 staged ONCE for both tiles)
//   warp 1        MMA issuer, per key block j:  O_A += P_A(j) V(j); S_A(j+1) = Q_A K(j+1)^T; O_B += P_B(j) V(j);
//                 S_B(j+1) = Q_B K(j+1)^T.  tcgen05.mma executes in issue order, so S_t(j+1) may overwrite the TMEM
//                 columns P_t(j) lives in right behind the PV that reads them.
//   warp 2        TMEM allocator (all 512 columns: S_A 0-127, S_B 128-255, O_A 256-383, O_B 384-511)
//   warps 4..11   softmax of tile A, warps 12..19 softmax of tile B.  TWO threads per query row (warps w and w+4 own the
//                 same 32 TMEM lanes; one takes key columns 0-63 of the block, the other 64-127): tcgen05.ld the half
//                 row, block max exchanged through shared memory + a 64-thread named barrier, running max with lazy
//                 rescale of O, exp2, bf16 P written over S, final O / l -> global.
// While the softmax warps of one tile work on block j, the tensor pipe runs the other tile's PV(j) + S(j+1).  One warp per
// scheduler only reaches about half of the MUFU rate (measured: r01c ncu capture, xu pipe 44 %), hence two warps per
// scheduler and tile: the MUFU pipe (16 ex2 / clk / SM = 1024 clk per 128x128 block) and the tensor pipe (2 x 512 clk per
// block) can both stay busy.
// Registers: 640 threads x 96 at launch; warps 0-3 shrink to 64 and the softmax warps grow to 104 (setmaxnreg).
constexpr int ATTN2_THREADS = 640;
constexpr int ATTN2_XCH_FLOATS = 2 * 2 * 2 * 128;       // [slot][tile][half][row]
constexpr int ATTN2_SMEM = TILE_BYTES * (2 + 2 * KV_STAGES) + 1024 + 256 + ATTN2_XCH_FLOATS * 4;
constexpr uint32_t TM2_S = 0, TM2_O = 256; // + 128 * tile

#ifdef SNB_ATTN_TRACE
// debug build only (SNB_NVCC_FLAGS=-DSNB_ATTN_TRACE): clock64() stamps of CTA (0,0,0): [role 0..4][block j < 16][event < 8]
__device__ long long g_attn_trace[6 * 16 * 8];
#define ATTN_TRACE(role, j, ev) do { if (trace_on && (j) < 16) g_attn_trace[((role) * 16 + (j)) * 8 + (ev)] = clock64(); } while (0)
#else
#define ATTN_TRACE(role, j, ev) do { } while (0)
#endif

template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                    // 2 tiles
    uint8_t *sK = smem + TILE_BYTES * 2;                   // KV_STAGES
    uint8_t *sV = smem + TILE_BYTES * (2 + KV_STAGES);     // KV_STAGES
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TILE_BYTES * (2 + 2 * KV_STAGES));
    uint64_t *q_full = bars;                               // [2] per tile
    uint64_t *k_full = bars + 2, *k_empty = bars + 4, *v_full = bars + 6, *v_empty = bars + 8;   // [2] per stage
    uint64_t *s_full = bars + 10, *p_ready = bars + 12, *pv_done = bars + 14;                    // [2] per tile
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
    float *xch = reinterpret_cast<float *>(bars + 32);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int q0 = blockIdx.x * (2 * BQ), head = blockIdx.y, env = blockIdx.z;
    const int n_tok = args.n_tok;
    const int n_kv = (n_tok + BKV - 1) / BKV;
    const bool has_b = q0 + BQ < n_tok;                    // the second tile holds at least one real query
#ifdef SNB_ATTN_TRACE
    const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp < 4 || (warp & 3) == 0);
    if (trace_on && warp == 0) {
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_attn_trace[5 * 128 + 0] = clock64(); g_attn_trace[5 * 128 + 3] = gt;
    }
#endif

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQKV);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&q_full[s], 1);
            tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1);
            tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
            tc::mbar_init(&s_full[s], 1); tc::mbar_init(&p_ready[s], 256); tc::mbar_init(&pv_done[s], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef SNB_ATTN_TRACE
    if (trace_on && warp == 0) g_attn_trace[5 * 128 + 1] = clock64();
#endif

    if (warp < 4) {
        setmaxnreg_dec<64>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            const int cq = head * HD, ck = 512 + head * HD, cv = 1024 + head * HD;
            for (int t = 0; t < (has_b ? 2 : 1); ++t) {
                tc::mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
                tc::tma_load_3d(sQ + t * TILE_BYTES, &tmQKV, &q_full[t], cq, q0 + t * BQ, env);
                tc::tma_load_3d(sQ + t * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &q_full[t], cq + 64, q0 + t * BQ, env);
            }
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                tc::mbar_wait(&k_empty[st], ph ^ 1);
                tc::mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
                tc::tma_load_3d(sK + st * TILE_BYTES, &tmQKV, &k_full[st], ck, j * BKV, env);
                tc::tma_load_3d(sK + st * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &k_full[st], ck + 64, j * BKV, env);
                tc::mbar_wait(&v_empty[st], ph ^ 1);
                tc::mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
                tc::tma_load_3d(sV + st * TILE_BYTES, &tmQKV, &v_full[st], cv, j * BKV, env);
                tc::tma_load_3d(sV + st * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &v_full[st], cv + 64, j * BKV, env);
            }
        } else if (warp == 1) {
          // ===================== MMA issuer =====================
          // tmem_base comes out of shared memory; the broadcast makes it (and every address derived from it) provably
          // warp-uniform, so the tcgen05.mma operands live in uniform registers instead of being converted per instruction
          const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
          if (tc::elect_one()) {
            auto kv_cols = [&](int j) { // keys of block j rounded up to the UMMA N granularity (16)
                const int rem = n_tok - j * BKV;
                return rem >= BKV ? BKV : ((rem + 15) & ~15);
            };
            // Issue cost matters (tools/mma_bench*.cu): building a shared-memory descriptor per tcgen05.mma costs ~130 clk per
            // instruction on the issuing thread, twice the 64 clk a 128x128x16 MMA runs.  All descriptors are therefore built
            // once; inside the unrolled K loops an MMA's operands are `base + compile-time constant`.
            const uint64_t qd0 = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024), qd1 = tc::make_smem_desc_sw128(tc::smem_u32(sQ + TILE_BYTES), 16, 1024);
            const uint64_t kd0 = tc::make_smem_desc_sw128(tc::smem_u32(sK), 16, 1024), kd1 = tc::make_smem_desc_sw128(tc::smem_u32(sK + TILE_BYTES), 16, 1024);
            // V is the MN-major B operand: 16 keys = 2 groups of 8 rows (SBO 1024 B); the two 64-wide head-dim boxes are LBO = 16 KB apart
            const uint64_t vd0 = tc::make_smem_desc_sw128(tc::smem_u32(sV), TILE_BYTES / 2, 1024), vd1 = tc::make_smem_desc_sw128(tc::smem_u32(sV + TILE_BYTES), TILE_BYTES / 2, 1024);
            auto issue_S = [&](int t, int j) {   // caller has waited for Q_t and K(j)
                const uint64_t qd = t ? qd1 : qd0, kd = (j & 1) ? kd1 : kd0;
                const uint32_t idesc = tc::make_idesc_bf16(BQ, (uint32_t)kv_cols(j), 0, 0);
                const uint32_t d = tmem_base + TM2_S + t * 128;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint64_t off = (uint64_t)(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
                    tc::umma_ss(d, qd + off, kd + off, idesc, k != 0 ? 1u : 0u);
                }
                tc::umma_commit(&s_full[t]);
            };
            constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, HD, 0, 1); // B = V is MN-major (head dim contiguous)
            auto issue_PV = [&](int t, int j) {  // caller has waited for P_t(j) and V(j)
                const uint64_t vd = (j & 1) ? vd1 : vd0;
                const uint32_t p_tmem = tmem_base + TM2_S + t * 128;
                const uint32_t o_tmem = tmem_base + TM2_O + t * 128;
                const int ksteps = kv_cols(j) / 16;
                if (ksteps == BKV / 16) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        tc::umma_ts(o_tmem, p_tmem + k * 8, vd + (uint64_t)((k * 2048) >> 4), idesc_pv, (j | k) != 0 ? 1u : 0u);
                } else {
                    for (int k = 0; k < ksteps; ++k)
                        tc::umma_ts(o_tmem, p_tmem + k * 8, vd + (uint64_t)((k * 2048) >> 4), idesc_pv, (j | k) != 0 ? 1u : 0u);
                }
                tc::umma_commit(&pv_done[t]);
            };
            tc::mbar_wait(&q_full[0], 0);
            tc::mbar_wait(&k_full[0], 0);
            tc::tc_fence_after();
            issue_S(0, 0);
            if (has_b) {
                tc::mbar_wait(&q_full[1], 0);
                tc::tc_fence_after();
                issue_S(1, 0);
            }
            tc::umma_commit(&k_empty[0]);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t ph = j & 1, ring_ph = (j >> 1) & 1;
                const bool more = j + 1 < n_kv;
                tc::mbar_wait(&p_ready[0], ph);
                ATTN_TRACE(4, j, 0);
                tc::mbar_wait(&v_full[st], ring_ph);
                tc::tc_fence_after();
                ATTN_TRACE(4, j, 1);
                issue_PV(0, j);
                if (more) {
                    tc::mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
                    tc::tc_fence_after();
                    issue_S(0, j + 1);
                }
                ATTN_TRACE(4, j, 2);
                if (has_b) {
                    tc::mbar_wait(&p_ready[1], ph);
                    tc::tc_fence_after();
                    ATTN_TRACE(4, j, 3);
                    issue_PV(1, j);
                    if (more) issue_S(1, j + 1);
                    ATTN_TRACE(4, j, 4);
                }
                tc::umma_commit(&v_empty[st]);
                if (more) tc::umma_commit(&k_empty[st ^ 1]);
            }
          }
        }
    } else {
        setmaxnreg_inc<104>();
        // ===================== softmax / correction / epilogue: tile t, column half `half` of each key block =====================
        const int t = (warp - 4) >> 3;
        const int half = ((warp - 4) >> 2) & 1;
        if (t == 0 || has_b) {
            const int quarter = warp & 3;
            const int row_in_tile = quarter * 32 + lane;
            const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
            const uint32_t s_addr = tmem_base + lane_addr + TM2_S + t * 128 + half * 64;   // fp32 scores of my 64 keys
            const uint32_t p_addr = tmem_base + lane_addr + TM2_S + t * 128 + half * 32;   // bf16 pairs of my 64 keys
            const uint32_t o_addr = tmem_base + lane_addr + TM2_O + t * 128 + half * 64;   // my 64 head-dim columns of O
            const int pair_bar = 1 + t * 4 + quarter;                                      // named barrier of warps w, w+4
            float *x_mine = xch + (t * 2 + half) * 128 + row_in_tile;
            float *x_peer = xch + (t * 2 + (half ^ 1)) * 128 + row_in_tile;
            const float c = args.scale_log2;
            float m_used = -INFINITY; // running max the exponents are taken against (raw score units)
            float l = 0.0f;           // partial row sum over my key columns
            for (int j = 0; j < n_kv; ++j) {
                const uint32_t ph = j & 1;
                ATTN_TRACE(t * 2 + half, j, 0);
                tc::mbar_wait(&s_full[t], ph);
                ATTN_TRACE(t * 2 + half, j, 1);
                __syncwarp();                 // reconverge before the .sync.aligned TMEM loads
                tc::tc_fence_after();
#if defined(SNB_ATTN_TRACE) && SNB_ATTN_TRACE >= 2   // debug: no softmax work at all -> MMA durations without any contention
                tc::tc_fence_before();
                tc::mbar_arrive(&p_ready[t]);
                continue;
#endif
                uint32_t s0[16], s1[16], s2[16], s3[16];   // x16 loads: 16-register operand groups are easier to place than x32
                tc::tmem_ld_32x16(s_addr, s0);
                tc::tmem_ld_32x16(s_addr + 16, s1);
                tc::tmem_ld_32x16(s_addr + 32, s2);
                tc::tmem_ld_32x16(s_addr + 48, s3);
                tc::tmem_ld_wait();
                ATTN_TRACE(t * 2 + half, j, 2);
                const int valid = n_tok - j * BKV - half * 64; // keys of my half block that exist
                if (valid < 64) {                              // ragged last block: keys that do not exist score -inf
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (i >= valid) s0[i] = 0xff800000u;
                        if (16 + i >= valid) s1[i] = 0xff800000u;
                        if (32 + i >= valid) s2[i] = 0xff800000u;
                        if (48 + i >= valid) s3[i] = 0xff800000u;
                    }
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s2[i]), __uint_as_float(s3[i])));
                }
                // block max of the whole row: exchange with the thread that owns the other 64 keys.  Slots alternate with j,
                // so slot (j & 1) is rewritten only after the pair barrier of block j+1, which the peer passes after this read.
                // The barrier also orders the peer's S loads before my P stores (P of keys 64-127 overlays S columns 32-63).
                x_mine[ph * 512] = fmaxf(mx0, mx1);
                tc::named_bar_sync(pair_bar, 64);
                const float bmax = fmaxf(fmaxf(mx0, mx1), x_peer[ph * 512]);
                ATTN_TRACE(t * 2 + half, j, 3);
                if (j == 0) {
                    m_used = bmax;
                } else if (__any_sync(0xffffffffu, (bmax - m_used) * c > RESCALE_THRESHOLD)) {
                    // Refresh the running max.  The decision is uniform over the warp (tcgen05.ld/st are .sync.aligned) and
                    // identical in the peer warp (same rows, same bmax); lanes whose own max did not grow rescale by exactly 1.
                    tc::mbar_wait(&pv_done[t], (j - 1) & 1);              // O is stable once the previous PV has landed
                    __syncwarp();
                    tc::tc_fence_after();
                    const float m_new = fmaxf(m_used, bmax);
                    const float f = tc::ex2_approx((m_used - m_new) * c);
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t r[16];
                        tc::tmem_ld_32x16(o_addr + ch * 16, r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                        tc::tmem_st_32x16(o_addr + ch * 16, r);
                    }
                    l *= f;
                    m_used = m_new;
                }
                const float mc = m_used * c;
                float sum0 = 0.0f, sum1 = 0.0f;
                auto exp_pack = [&](const uint32_t (&sa)[16], const uint32_t (&sb)[16], uint32_t col) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float a0 = tc::ex2_approx(fmaf(__uint_as_float(sa[2 * i]), c, -mc)), a1 = tc::ex2_approx(fmaf(__uint_as_float(sa[2 * i + 1]), c, -mc));
                        const float b0 = tc::ex2_approx(fmaf(__uint_as_float(sb[2 * i]), c, -mc)), b1 = tc::ex2_approx(fmaf(__uint_as_float(sb[2 * i + 1]), c, -mc));
                        sum0 += a0 + a1; sum1 += b0 + b1;
                        pk[i] = tc::pack_bf16(a0, a1); pk[8 + i] = tc::pack_bf16(b0, b1);
                    }
                    tc::tmem_st_32x16(p_addr + col, pk);
                };
                exp_pack(s0, s1, 0);       // P columns [0,16) of my half  <- S columns [0,32)
                exp_pack(s2, s3, 16);      // P columns [16,32) of my half <- S columns [32,64)
                l += sum0 + sum1;
                ATTN_TRACE(t * 2 + half, j, 4);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&p_ready[t]);
                ATTN_TRACE(t * 2 + half, j, 5);
            }
            // final: O / l -> global (l = my partial sum + the peer's)
            x_mine[(n_kv & 1) * 512] = l;
            tc::named_bar_sync(pair_bar, 64);
            const float inv_l = 1.0f / (l + x_peer[(n_kv & 1) * 512]);
            tc::mbar_wait(&pv_done[t], (n_kv - 1) & 1);
            __syncwarp();
            tc::tc_fence_after();
            const int row = q0 + t * BQ + row_in_tile;
            bf16 *dst = args.out + ((size_t)env * n_tok + row) * (NHEAD * HD) + head * HD + half * 64;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t r[16];
                tc::tmem_ld_32x16(o_addr + ch * 16, r);
                tc::tmem_ld_wait();
                if (row < n_tok) {
                    uint4 *o4 = reinterpret_cast<uint4 *>(dst + ch * 16);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint4 o;
                        o.x = tc::pack_bf16(__uint_as_float(r[8 * q + 0]) * inv_l, __uint_as_float(r[8 * q + 1]) * inv_l);
                        o.y = tc::pack_bf16(__uint_as_float(r[8 * q + 2]) * inv_l, __uint_as_float(r[8 * q + 3]) * inv_l);
                        o.z = tc::pack_bf16(__uint_as_float(r[8 * q + 4]) * inv_l, __uint_as_float(r[8 * q + 5]) * inv_l);
                        o.w = tc::pack_bf16(__uint_as_float(r[8 * q + 6]) * inv_l, __uint_as_float(r[8 * q + 7]) * inv_l);
                        o4[q] = o;
                    }
                }
            }
            tc::tc_fence_before();
        }
    }
    __syncthreads();
#ifdef SNB_ATTN_TRACE
    if (trace_on && warp == 0) {
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_attn_trace[5 * 128 + 2] = clock64(); g_attn_trace[5 * 128 + 4] = gt;
    }
#endif
    if (warp == 2) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// iMID: R independent sequences of T <= 32 tokens (TransformerConcatLinear, diffusion.py:147).  One warp per
// (sequence, head); 8x8 scores are far below tensor-core tile sizes, so this runs on the CUDA cores.
__global__ void attn_small_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int n_seq, int T, float scale)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= n_seq * NHEAD) return;
    const int seq = gw / NHEAD, head = gw % NHEAD;
    const bf16 *base = qkv + (size_t)seq * T * 1536 + head * HD;
    // lane owns head-dim elements [4*lane, 4*lane+4)
    for (int i = 0; i < T; ++i) {
        float q[4];
        const uint2 qv = *reinterpret_cast<const uint2 *>(base + (size_t)i * 1536 + 4 * lane);
        const __nv_bfloat162 qa = *reinterpret_cast<const __nv_bfloat162 *>(&qv.x), qb = *reinterpret_cast<const __nv_bfloat162 *>(&qv.y);
        q[0] = __bfloat162float(qa.x); q[1] = __bfloat162float(qa.y); q[2] = __bfloat162float(qb.x); q[3] = __bfloat162float(qb.y);
        float sc[32];
        float mx = -INFINITY;
        for (int j = 0; j < T; ++j) {
            const uint2 kv = *reinterpret_cast<const uint2 *>(base + 512 + (size_t)j * 1536 + 4 * lane);
            const __nv_bfloat162 ka = *reinterpret_cast<const __nv_bfloat162 *>(&kv.x), kb = *reinterpret_cast<const __nv_bfloat162 *>(&kv.y);
            float d = q[0] * __bfloat162float(ka.x) + q[1] * __bfloat162float(ka.y) + q[2] * __bfloat162float(kb.x) + q[3] * __bfloat162float(kb.y);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            sc[j] = d * scale;
            mx = fmaxf(mx, sc[j]);
        }
        float den = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < T; ++j) {
            const float p = __expf(sc[j] - mx);
            den += p;
            const uint2 vv = *reinterpret_cast<const uint2 *>(base + 1024 + (size_t)j * 1536 + 4 * lane);
            const __nv_bfloat162 va = *reinterpret_cast<const __nv_bfloat162 *>(&vv.x), vb = *reinterpret_cast<const __nv_bfloat162 *>(&vv.y);
            acc[0] += p * __bfloat162float(va.x); acc[1] += p * __bfloat162float(va.y);
            acc[2] += p * __bfloat162float(vb.x); acc[3] += p * __bfloat162float(vb.y);
        }
        const float inv = 1.0f / den;
        uint2 o;
        o.x = tc::pack_bf16(acc[0] * inv, acc[1] * inv);
        o.y = tc::pack_bf16(acc[2] * inv, acc[3] * inv);
        *reinterpret_cast<uint2 *>(out + ((size_t)seq * T + i) * 512 + head * HD + 4 * lane) = o;
    }
}

} // namespace

int snb_attn_plan(AttnPlan *plan, const bf16 *qkv, int n_env, int n_tok)
{
    SNB_REQUIRE(n_env > 0 && n_tok > 0, SNB_EINVAL, "attention: bad sizes");
    plan->n_env = n_env; plan->n_tok = n_tok;
    return snb_make_tmap_3d(&plan->tmQKV, qkv, (uint64_t)n_env, (uint64_t)n_tok, 1536, BKV);
}

int snb_attn_launch(const AttnPlan *plan, bf16 *out, cudaStream_t stream)
{
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static bool use_v1 = false;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM);
        if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN2_SMEM);
        const char *e = getenv("SNB_ATTN_V1");
        use_v1 = e && e[0] == '1';
    });
    SNB_CUDA_TRY(attr_err);
    AttnArgs a;
    a.out = out; a.n_tok = plan->n_tok;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
    if (use_v1) {
        dim3 grid((plan->n_tok + BQ - 1) / BQ, NHEAD, plan->n_env);
        attn_fwd_kernel<<<grid, ATTN_THREADS, ATTN_SMEM, stream>>>(plan->tmQKV, a);
    } else {
        dim3 grid((plan->n_tok + 2 * BQ - 1) / (2 * BQ), NHEAD, plan->n_env);
        attn_fwd2_kernel<<<grid, ATTN2_THREADS, ATTN2_SMEM, stream>>>(plan->tmQKV, a);
    }
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

#ifdef SNB_ATTN_TRACE
extern "C" int snb_debug_attn_trace(long long *out)
{
    return cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(long long) * 6 * 16 * 8) == cudaSuccess ? 0 : -1;
}
#endif

int snb_attn_small_launch(const bf16 *qkv, bf16 *out, int n_seq, int T, cudaStream_t stream)
{
    SNB_REQUIRE(T >= 1 && T <= 32, SNB_EUNSUPPORTED, "attention(iMID): T=%d unsupported", T);
    const int warps = n_seq * NHEAD;
    const int threads = 256;
    attn_small_kernel<<<(warps * 32 + threads - 1) / threads, threads, 0, stream>>>(qkv, out, n_seq, T, 1.0f / sqrtf((float)HD));
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
