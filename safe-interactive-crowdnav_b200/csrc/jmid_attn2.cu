// csrc/jmid_attn2.cu -- second-generation flash attention of the JMID noise network (one unmasked sequence of T*A*S tokens per
// environment, models/diffusion.py:196-204), tcgen05 + TMEM + TMA, persistent.
//
// What the clock64 trace of the first kernel (jmid_attn.cu; profiles/r02_attn_trace_r01_kernel_chunk512.txt) showed per 128-key
// block: 2930 clk, of which the tensor pipe was busy 2048; the softmax warps' chain was  wait S 194 + tcgen05.ld 166 + max / exp
// 1992 + P -> shared memory 471 + arrive 158, and the MMA issuer stalled ~550 clk per block on the 3-slot K / V ring (one 32 KB
// tile of prefetch distance < L2 latency + transfer).  Both come from P living in shared memory: it cost 64 KB of the 227 KB (hence
// the shallow ring) and 512 clk of LSU store bandwidth per block.  Here:
//   * 64-key blocks.  S_t is 64 TMEM columns, so the 512 columns hold O_A, O_B (2 x 128), S_A, S_B (2 x 64) AND a double-buffered
//     bf16 P_A, P_B (4 x 32): P never touches shared memory.  The softmax thread stores its P row with one tcgen05.st (32 registers)
//     and  O_t += P_t V  is a TS MMA (A operand from TMEM).
//   * shared memory = Q_A, Q_B (64 KB) + a TEN-slot ring of 16 KB K / V blocks (160 KB): the producer runs five blocks ahead.
//   * row maximum with FMNMX3 (3-input max, sm_100), half the instructions of the 2-input tree.
// Roles as before: warp 0 TMA producer, warp 1 MMA issuer (one elected thread, precomputed descriptors), warp 2 TMEM allocator,
// warps 4-7 / 8-11 softmax + epilogue of tile A / B, one thread per query row, lazy rescale of O, 1 exponential in 4 on the FMA pipe.
#include <cfloat>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "jmid_internal.h"
#include "tc_utils.cuh"

namespace {

constexpr int HD = 128, NHEAD = 4, BQ = 128, BK = 64;
constexpr int Q_BYTES = BQ * HD * 2;      // 32 KB: two 64-column boxes of 16 KB
constexpr int KV_BYTES = BK * HD * 2;     // 16 KB: two 64-column boxes of 8 KB
constexpr int RING = 10;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_O = 0, TM_S = 256, TM_P = 384;   // O_t: +128 t; S_t: +64 t; P_t[b]: +64 t + 32 b
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr int THREADS = 384;
constexpr int SMEM = Q_BYTES * 2 + KV_BYTES * RING + 1024 + 512;
#ifndef SNB_ATTN_POLY_EVERY
#define SNB_ATTN_POLY_EVERY 4
#endif

struct Args {
    bf16 *out;
    int n_tok, n_items, n_qp;
    float scale_log2;
};

#ifdef SNB_ATTN_TRACE
// debug build only (SNB_NVCC_FLAGS=-DSNB_ATTN_TRACE): clock64() stamps of CTA 0's SECOND work item: [role A, B, MMA][block j < 32][event < 8]
__device__ long long g_attn2_trace[3 * 32 * 8 + 8];
#define TR(role, j, ev) do { if (trace_on && w == (int)(blockIdx.x + gridDim.x) && (j) < 32) g_attn2_trace[((role) * 32 + (j)) * 8 + (ev)] = clock64(); } while (0)
#else
#define TR(role, j, ev) do { } while (0)
#endif

template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__global__ void __launch_bounds__(THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Args args)
{
    constexpr int POLY_EVERY = SNB_ATTN_POLY_EVERY;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // 2 tiles of 128 queries
    uint8_t *sKV = smem + Q_BYTES * 2;                    // ring of RING blocks of 64 keys (K or V)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Q_BYTES * 2 + KV_BYTES * RING);
    uint64_t *q_full = bars, *q_empty = bars + 2, *o_free = bars + 4, *s_full = bars + 6, *s_free = bars + 8;
    // [tile][P buffer].  p_ready is per BUFFER: the softmax warps may finish iteration G + 1 (S(G+1) is issued before the issuer waits
    // for P(G)) while the issuer is still held up before P(G) -- a per-tile barrier would then be two phases ahead of its waiter.
    uint64_t *p_ready = bars + 10, *pv_done = bars + 14;
    uint64_t *kv_full = bars + 18, *kv_empty = bars + 18 + RING;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 18 + 2 * RING);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_tok = args.n_tok;
    const int n_kv = (n_tok + BK - 1) / BK;
    auto decode = [&](int w, int &q0, int &head, int &env, bool &has_b) {
        const int qp = w % args.n_qp, eh = w / args.n_qp;
        q0 = qp * (2 * BQ); head = eh % NHEAD; env = eh / NHEAD;
        has_b = q0 + BQ < n_tok;
    };

#ifdef SNB_ATTN_TRACE
    const bool trace_on = blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 4 || warp == 8);
#endif
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmKV);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1); tc::mbar_init(&o_free[s], 128);
            tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_free[s], 128);
            tc::mbar_init(&p_ready[2 * s], 128); tc::mbar_init(&p_ready[2 * s + 1], 128);
            tc::mbar_init(&pv_done[2 * s], 1); tc::mbar_init(&pv_done[2 * s + 1], 1);
        }
        for (int s = 0; s < RING; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    if (warp < 4) {
        setmaxnreg_dec<88>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            int c = 0;
            int it[2] = {0, 0};
            for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
                int q0, head, env; bool has_b;
                decode(w, q0, head, env, has_b);
                const int cq = head * HD, ck = 512 + head * HD, cv = 1024 + head * HD;
                for (int t = 0; t < (has_b ? 2 : 1); ++t) {
                    tc::mbar_wait(&q_empty[t], (it[t] & 1) ^ 1);
                    ++it[t];
                    tc::mbar_arrive_expect_tx(&q_full[t], Q_BYTES);
                    tc::tma_load_3d(sQ + t * Q_BYTES, &tmQ, &q_full[t], cq, q0 + t * BQ, env);
                    tc::tma_load_3d(sQ + t * Q_BYTES + Q_BYTES / 2, &tmQ, &q_full[t], cq + 64, q0 + t * BQ, env);
                }
                auto load = [&](int col, int j) {
                    const int slot = c % RING;
                    const uint32_t ph = (c / RING) & 1;
                    ++c;
                    tc::mbar_wait(&kv_empty[slot], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&kv_full[slot], KV_BYTES);
                    tc::tma_load_3d(sKV + slot * KV_BYTES, &tmKV, &kv_full[slot], col, j * BK, env);
                    tc::tma_load_3d(sKV + slot * KV_BYTES + KV_BYTES / 2, &tmKV, &kv_full[slot], col + 64, j * BK, env);
                };
                load(ck, 0);
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) load(ck, j + 1);
                    load(cv, j);
                }
            }
        } else if (warp == 1) {
          // ===================== MMA issuer =====================
          const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
          if (tc::elect_one()) {
            auto kv_cols = [&](int j) {
                const int rem = n_tok - j * BK;
                return rem >= BK ? BK : ((rem + 15) & ~15);
            };
            const uint64_t qd0 = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);
            const uint64_t kd0 = tc::make_smem_desc_sw128(tc::smem_u32(sKV), 16, 1024);               // K: K-major B operand
            const uint64_t vd0 = tc::make_smem_desc_sw128(tc::smem_u32(sKV), KV_BYTES / 2, 1024);     // V: MN-major, 64-dim boxes 8 KB apart
            constexpr uint64_t Q_DESC = Q_BYTES >> 4, KV_DESC = KV_BYTES >> 4;
            auto issue_S = [&](int t, int slot, int j) {
                const uint64_t qd = qd0 + (uint64_t)t * Q_DESC, kd = kd0 + (uint64_t)slot * KV_DESC;
                const uint32_t idesc = tc::make_idesc_bf16(BQ, (uint32_t)kv_cols(j), 0, 0);
                const uint32_t d = tmem_base + TM_S + t * 64;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    tc::umma_ss(d, qd + (uint64_t)(((k >> 2) * (Q_BYTES / 2) + (k & 3) * 32) >> 4),
                                kd + (uint64_t)(((k >> 2) * (KV_BYTES / 2) + (k & 3) * 32) >> 4), idesc, k != 0 ? 1u : 0u);
                tc::umma_commit(&s_full[t]);
            };
            constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, HD, 0, 1);   // A = P from TMEM, B = V MN-major
            auto issue_PV = [&](int t, int slot, int j, int b) {
                const uint64_t vd = vd0 + (uint64_t)slot * KV_DESC;
                const uint32_t o_tmem = tmem_base + TM_O + t * 128, p_tmem = tmem_base + TM_P + t * 64 + b * 32;
                const int ksteps = kv_cols(j) / 16;
                if (ksteps == BK / 16) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc::umma_ts(o_tmem, p_tmem + k * 8, vd + (uint64_t)((k * 2048) >> 4), idesc_pv, (j | k) != 0 ? 1u : 0u);
                } else {
                    for (int k = 0; k < ksteps; ++k)
                        tc::umma_ts(o_tmem, p_tmem + k * 8, vd + (uint64_t)((k * 2048) >> 4), idesc_pv, (j | k) != 0 ? 1u : 0u);
                }
                tc::umma_commit(&pv_done[2 * t + b]);
            };
            int c = 0, slot;
            auto next_tile = [&]() {
                slot = c % RING;
                const uint32_t ph = (c / RING) & 1;
                ++c;
                tc::mbar_wait(&kv_full[slot], ph);
            };
            int g[2] = {0, 0}, it[2] = {0, 0};
            for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
                int q0, head, env; bool has_b;
                decode(w, q0, head, env, has_b);
                const int nt = has_b ? 2 : 1;
                next_tile();                             // K0
                for (int t = 0; t < nt; ++t) {
                    tc::mbar_wait(&q_full[t], it[t] & 1);
                    if (g[t] > 0) tc::mbar_wait(&s_free[t], (g[t] - 1) & 1);
                    tc::tc_fence_after();
                    issue_S(t, slot, 0);
                    if (n_kv == 1) tc::umma_commit(&q_empty[t]);
                }
                tc::umma_commit(&kv_empty[slot]);
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) {
                        next_tile();                     // K(j+1)
                        TR(2, j, 0);
                        for (int t = 0; t < nt; ++t) {
                            tc::mbar_wait(&s_free[t], (g[t] + j) & 1);
                            tc::tc_fence_after();
                            if (t == 0) TR(2, j, 1);
                            issue_S(t, slot, j + 1);
                            if (j + 2 == n_kv) tc::umma_commit(&q_empty[t]);
                        }
                        tc::umma_commit(&kv_empty[slot]);
                        TR(2, j, 2);
                    }
                    next_tile();                         // V(j)
                    TR(2, j, 3);
                    for (int t = 0; t < nt; ++t) {
                        if (j == 0) tc::mbar_wait(&o_free[t], (it[t] & 1) ^ 1);
                        const int G = g[t] + j;
                        tc::mbar_wait(&p_ready[2 * t + (G & 1)], (G >> 1) & 1);
                        tc::tc_fence_after();
                        TR(2, j, 4 + 2 * t);
                        issue_PV(t, slot, j, G & 1);
                        TR(2, j, 5 + 2 * t);
                    }
                    tc::umma_commit(&kv_empty[slot]);
                }
                for (int t = 0; t < nt; ++t) { g[t] += n_kv; ++it[t]; }
            }
          }
        }
    } else {
        setmaxnreg_inc<208>();
        // ===================== softmax / correction / epilogue of tile t =====================
        const int t = (warp - 4) >> 2;
        const uint32_t tmem_base = *tmem_slot;
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const uint32_t s_addr = tmem_base + lane_addr + TM_S + t * 64;
        const uint32_t o_addr = tmem_base + lane_addr + TM_O + t * 128;
        const uint32_t p_addr = tmem_base + lane_addr + TM_P + t * 64;
        const float c = args.scale_log2;
        int g = 0;
        // PV(G) (G = running iteration of this tile) used P buffer G & 1 for the (G >> 1)-th time: its completion is phase (G >> 1) & 1
        auto wait_pv = [&](int G) { tc::mbar_wait(&pv_done[2 * t + (G & 1)], (G >> 1) & 1); };
        for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
            int q0, head, env; bool has_b;
            decode(w, q0, head, env, has_b);
            if (t == 1 && !has_b) continue;
            float m_used = -INFINITY, l = 0.0f;
            // one 64-key block.  RAGGED is a compile-time flag: the per-element masking of keys that do not exist (64 ISETP + 64 SEL,
            // a quarter of the loop's instructions when ptxas if-converts it) is only compiled into the copy that runs the LAST block.
            auto block = [&](int j, auto ragged_tag) {
                constexpr bool RAGGED = decltype(ragged_tag)::value;
                const int G = g + j;
                TR(t, j, 0);
                tc::mbar_wait(&s_full[t], G & 1);
                TR(t, j, 1);
                __syncwarp();
                tc::tc_fence_after();
                uint32_t s0[32], s1[32];
                tc::tmem_ld_32x32(s_addr, s0);
                tc::tmem_ld_32x32(s_addr + 32, s1);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&s_free[t]);
                TR(t, j, 2);
                if constexpr (RAGGED) {
                    const int valid = n_tok - j * BK;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= valid) s0[i] = 0xff800000u;
                        if (32 + i >= valid) s1[i] = 0xff800000u;
                    }
                }
                float mx0 = __uint_as_float(s0[0]), mx1 = __uint_as_float(s1[0]);
#pragma unroll
                for (int i = 1; i < 31; i += 2) {
                    mx0 = fmax3(mx0, __uint_as_float(s0[i]), __uint_as_float(s0[i + 1]));
                    mx1 = fmax3(mx1, __uint_as_float(s1[i]), __uint_as_float(s1[i + 1]));
                }
                const float bmax = fmax3(mx0, mx1, fmaxf(__uint_as_float(s0[31]), __uint_as_float(s1[31])));
                if (j == 0) {
                    m_used = bmax;
                } else if (__any_sync(0xffffffffu, (bmax - m_used) * c > RESCALE_THRESHOLD)) {
                    wait_pv(G - 1);                                       // O_t is stable once PV_t(G-1) has landed
                    __syncwarp();
                    tc::tc_fence_after();
                    const float m_new = fmaxf(m_used, bmax);
                    const float f = tc::ex2_approx((m_used - m_new) * c);
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t r[16];
                        tc::tmem_ld_32x16(o_addr + ch * 16, r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                        tc::tmem_st_32x16(o_addr + ch * 16, r);
                    }
                    tc::tmem_st_wait();
                    tc::tc_fence_before();
                    l *= f;
                    m_used = m_new;
                }
                const float mc = m_used * c;
                const uint64_t c2 = tc::f2_pack(c, c), nmc2 = tc::f2_pack(-mc, -mc);
                uint64_t sumA = tc::f2_pack(0.0f, 0.0f), sumB = sumA;
                uint32_t p[32];
                auto exp_pack = [&](const uint32_t (&s)[32], int base) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint64_t x2 = tc::f2_fma(tc::f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), c2, nmc2);
                        uint64_t e2;
                        if (POLY_EVERY > 0 && (i % POLY_EVERY) == POLY_EVERY - 1) {
                            e2 = tc::f2_exp2_poly(x2);
                        } else {
                            float xl, xh;
                            tc::f2_unpack(x2, xl, xh);
                            e2 = tc::f2_pack(tc::ex2_approx(xl), tc::ex2_approx(xh));
                        }
                        if (i & 1) sumB = tc::f2_add(sumB, e2); else sumA = tc::f2_add(sumA, e2);
                        float el, eh;
                        tc::f2_unpack(e2, el, eh);
                        p[base + i] = tc::pack_bf16(el, eh);
                    }
                };
                exp_pack(s0, 0); exp_pack(s1, 16);
                {
                    float a0, a1, b0, b1;
                    tc::f2_unpack(sumA, a0, a1);
                    tc::f2_unpack(sumB, b0, b1);
                    l += (a0 + a1) + (b0 + b1);
                }
                TR(t, j, 3);
                if (G >= 2) wait_pv(G - 2);                               // the PV that read this P buffer two iterations ago
                TR(t, j, 4);
                __syncwarp();
                tc::tc_fence_after();
                tc::tmem_st_32x32(p_addr + (G & 1) * 32, p);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&p_ready[2 * t + (G & 1)]);
                TR(t, j, 5);
            };
            const bool last_ragged = (n_tok % BK) != 0;
            for (int j = 0; j < n_kv - 1; ++j) block(j, std::false_type{});
            if (last_ragged) block(n_kv - 1, std::true_type{}); else block(n_kv - 1, std::false_type{});
            // final: O / l -> global
            TR(t, 31, 0);
            wait_pv(g + n_kv - 1);
            TR(t, 31, 1);
            __syncwarp();
            tc::tc_fence_after();
            const float inv_l = 1.0f / l;
            const int row = q0 + t * BQ + row_in_tile;
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
            tc::mbar_arrive(&o_free[t]);
            TR(t, 31, 2);
            g += n_kv;
        }
    }
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<TMEM_COLS>(__shfl_sync(0xffffffffu, *tmem_slot, 0));
}

} // namespace

int snb_attn2_plan(Attn2Plan *plan, const bf16 *qkv, int n_env, int n_tok)
{
    SNB_REQUIRE(n_env > 0 && n_tok > 0, SNB_EINVAL, "attention: bad sizes");
    plan->n_env = n_env; plan->n_tok = n_tok;
    int rc = snb_make_tmap_3d(&plan->tmQ, qkv, (uint64_t)n_env, (uint64_t)n_tok, 1536, BQ);
    if (rc) return rc;
    return snb_make_tmap_3d(&plan->tmKV, qkv, (uint64_t)n_env, (uint64_t)n_tok, 1536, BK);
}

int snb_attn2_launch(const Attn2Plan *plan, bf16 *out, cudaStream_t stream)
{
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(attn2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); });
    SNB_CUDA_TRY(attr_err);
    Args a;
    a.out = out; a.n_tok = plan->n_tok;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
    a.n_qp = (plan->n_tok + 2 * BQ - 1) / (2 * BQ);
    a.n_items = a.n_qp * NHEAD * plan->n_env;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SNB_CUDA_TRY(cudaGetDevice(&dev));
        SNB_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = a.n_items < num_sms ? a.n_items : num_sms;
    attn2_fwd_kernel<<<grid, THREADS, SMEM, stream>>>(plan->tmQ, plan->tmKV, a);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

#ifdef SNB_ATTN_TRACE
extern "C" int snb_debug_attn2_trace(long long *out)
{
    return cudaMemcpyFromSymbol(out, g_attn2_trace, sizeof(long long) * (3 * 32 * 8 + 8)) == cudaSuccess ? 0 : -1;
}
#endif
