// csrc/jmid_attn.cu -- fused multi-head self-attention of the JMID noise network on tcgen05.
//
// The reference flattens the T*A*S tokens of one environment into ONE unmasked sequence and runs
// nn.MultiheadAttention(d=512, heads=4) over it (models/diffusion.py:196-204, quirk q1).  This kernel computes, per
// (environment, head, 128-query tile):  O = softmax(Q K^T / sqrt(128)) V  flash-style, never materialising the
// N x N score matrix in HBM.  Structure and measurements: see the comment above attn_fwd_kernel.
#include <cfloat>
#include <mutex>
#include <type_traits>

#include "jmid_internal.h"
#include "tc_utils.cuh"

namespace {

constexpr int HD = 128;       // head dim
constexpr int NHEAD = 4;
constexpr int BQ = 128;       // queries per tile (two tiles per CTA)
constexpr int BKV = 128;      // keys per block
constexpr int TILE_BYTES = BQ * HD * 2; // 32 KB: two 64-column boxes of 16 KB
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TM_S = 0, TM_O = 256;  // + 128 * tile
constexpr float RESCALE_THRESHOLD = 8.0f; // log2 units: P <= 2^8 before the running max is refreshed

struct AttnArgs {
    bf16 *out;        // [n_env * n_tok, 512]
    int n_tok;
    int n_items;      // n_env * NHEAD * n_qp work items (environment, head, pair of 128-query tiles)
    int n_qp;         // query-tile pairs per (environment, head)
    float scale_log2; // log2(e) / sqrt(HD)
};

#ifdef SNB_ATTN_TRACE
// debug build only (SNB_NVCC_FLAGS=-DSNB_ATTN_TRACE): clock64() stamps of CTA (0,0,0): [role 0..4][block j < 16][event < 8]
__device__ long long g_attn_trace[6 * 16 * 8];
// stamps the SECOND work item of CTA 0 (steady state: its prologue overlaps the first item)
#define ATTN_TRACE(role, j, ev) do { if (trace_on && w == (int)(blockIdx.x + gridDim.x) && (j) < 16) g_attn_trace[((role) * 16 + (j)) * 8 + (ev)] = clock64(); } while (0)
#else
#define ATTN_TRACE(role, j, ev) do { } while (0)
#endif

__device__ __forceinline__ float fmax3(float a, float b, float c)   // FMNMX3: one instruction for two comparisons (sm_100)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------------------------------
// attn_fwd_kernel: softmax decoupled from the tensor pipe.  Two 128-query tiles (A, B) of one (environment, head) per CTA.
//   warp 0        TMA producer: Q_A, Q_B, then K / V blocks of 128 keys through ONE 3-slot ring in consumption order
//                 K0, K1, V0, K2, V1, ...
//   warp 1        MMA issuer.  S_t = Q_t K^T (SS) into TMEM; O_t += P_t V (SS: P_t comes from SHARED memory).  Because P does
//                 not overlay S, S_t(j+1) is issued as soon as the softmax warps have pulled S_t(j) into registers -- long
//                 before P_t(j) exists -- so the softmax warps never wait for the tensor pipe and the only serial chain left
//                 is softmax(j) -> softmax(j+1) (measured with the clock64 trace: the aliased-P variant spent 40 % of every
//                 block waiting for S).
//   warp 2        TMEM allocator (S_A 0-127, S_B 128-255, O_A 256-383, O_B 384-511)
//   warps 4..7    softmax of tile A, warps 8..11 of tile B; one thread per query row (no shuffles): tcgen05.ld the S row,
//                 release S, running max with lazy rescale of O, exp2, bf16 P row -> 128B-swizzled shared memory
//                 (the K-major A operand layout TMA would have produced), final O / l -> global.
// Registers: 384 threads x 168 at launch; warps 0-3 shrink to 88, the softmax warps grow to 208 (setmaxnreg).
#ifndef SNB_ATTN_POLY_EVERY
#define SNB_ATTN_POLY_EVERY 4
#endif
#ifndef SNB_ATTN_PV_WAIT_AFTER
#define SNB_ATTN_PV_WAIT_AFTER 1
#endif
#ifndef SNB_ATTN_LAG_AT
#define SNB_ATTN_LAG_AT 1      // tile B starts when tile A has finished this many quarters (+1) of its first block's exponentials
#endif
constexpr int ATTN_THREADS = 384;
constexpr int ATTN_RING = 3;
constexpr int ATTN_SMEM = TILE_BYTES * (2 + 2 + ATTN_RING) + 1024 + 256;

__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut, const AttnArgs args)
{
    constexpr int POLY_EVERY = SNB_ATTN_POLY_EVERY;   // one key pair in POLY_EVERY gets its 2^x from the FMA pipe (0 = none)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                    // 2 tiles
    uint8_t *sP = smem + TILE_BYTES * 2;                   // 2 tiles
    uint8_t *sKV = smem + TILE_BYTES * 4;                  // ring of ATTN_RING tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TILE_BYTES * (4 + ATTN_RING));
    uint64_t *q_full = bars;                               // [2] per tile
    uint64_t *kv_full = bars + 2, *kv_empty = bars + 5;    // [3] per ring slot
    uint64_t *s_full = bars + 8, *s_free = bars + 10, *p_ready = bars + 12, *pv_done = bars + 14;   // [2] per tile
    uint64_t *q_empty = bars + 16, *o_free = bars + 18;    // [2] per tile: Q_t may be refilled / O_t may be overwritten (next work item)
    uint64_t *lag = bars + 20;                             // tile A -> tile B: "half of my first block is done" (de-phases the two softmaxes)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 21);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int n_tok = args.n_tok;
    const int n_kv = (n_tok + BKV - 1) / BKV;
    // PERSISTENT: this CTA walks the work items blockIdx.x, blockIdx.x + gridDim.x, ...  Every role runs the same loop and decodes an
    // item the same way; all mbarrier phases come from running counters (iterations g_t = items x n_kv, items i_t, ring index c),
    // so the next item's Q / K0 are in flight and its S(0) is issued while the softmax warps still drain O of the current one.
    auto decode = [&](int w, int &q0, int &head, int &env, bool &has_b) {
        const int qp = w % args.n_qp, eh = w / args.n_qp;
        q0 = qp * (2 * BQ); head = eh % NHEAD; env = eh / NHEAD;
        has_b = q0 + BQ < n_tok;                           // the second tile holds at least one real query
    };
#ifdef SNB_ATTN_TRACE
    const bool trace_on = blockIdx.x == 0 && lane == 0 && (warp < 4 || (warp & 3) == 0);
    if (trace_on && warp == 0) {
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_attn_trace[5 * 128 + 0] = clock64(); g_attn_trace[5 * 128 + 3] = gt;
    }
#endif

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQKV);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1); tc::mbar_init(&o_free[s], 128);
            tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_free[s], 128); tc::mbar_init(&p_ready[s], 128); tc::mbar_init(&pv_done[s], 1);
        }
        tc::mbar_init(lag, 128);
        for (int s = 0; s < ATTN_RING; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 2); }   // kv_empty: both issuers
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
#ifdef SNB_ATTN_TRACE
    if (trace_on && warp == 0) g_attn_trace[5 * 128 + 1] = clock64();
#endif

    if (warp < 4) {
        setmaxnreg_dec<120>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            int c = 0;                                   // running index into the K0, K1, V0, K2, V1, ... sequence (all items)
            int it[2] = {0, 0};                          // items tile t has taken part in
            for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
                int q0, head, env; bool has_b;
                decode(w, q0, head, env, has_b);
                const int cq = head * HD, ck = 512 + head * HD, cv = 1024 + head * HD;
                for (int t = 0; t < (has_b ? 2 : 1); ++t) {
                    tc::mbar_wait(&q_empty[t], (it[t] & 1) ^ 1);     // the previous item's last S_t has read Q_t
                    ++it[t];
                    tc::mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
                    tc::tma_load_3d(sQ + t * TILE_BYTES, &tmQKV, &q_full[t], cq, q0 + t * BQ, env);
                    tc::tma_load_3d(sQ + t * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &q_full[t], cq + 64, q0 + t * BQ, env);
                }
                auto load = [&](int col, int j) {
                    const int slot = c % ATTN_RING;
                    const uint32_t ph = (c / ATTN_RING) & 1;
                    ++c;
                    tc::mbar_wait(&kv_empty[slot], ph ^ 1);
                    tc::mbar_arrive_expect_tx(&kv_full[slot], TILE_BYTES);
                    tc::tma_load_3d(sKV + slot * TILE_BYTES, &tmQKV, &kv_full[slot], col, j * BKV, env);
                    tc::tma_load_3d(sKV + slot * TILE_BYTES + TILE_BYTES / 2, &tmQKV, &kv_full[slot], col + 64, j * BKV, env);
                };
                load(ck, 0);
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) load(ck, j + 1);
                    load(cv, j);
                }
            }
        } else if (warp == 1 || warp == 3) {
          // ===================== MMA issuers: warp 1 issues tile A's MMAs, warp 3 tile B's =====================
          // One issuer per tile: with a single in-order issuer S_A(j+1) queued behind the wait for P_B(j-1) and PV_A(j) behind the
          // wait for S_B(j)'s consumer, which tied the two tiles together and pulled the half-block offset set up by `lag` back to
          // ~500 clk within three blocks (trace of the de-phased kernel) -- the softmaxes then ran in phase again and shared the MUFU
          // pipe.  Independent issuers only meet in the K/V ring: a slot is released when BOTH have committed their MMAs on it.
          const int t = warp == 1 ? 0 : 1;
          // tmem_base comes out of shared memory; the broadcast makes it (and every address derived from it) provably
          // warp-uniform, so the tcgen05.mma operands live in uniform registers instead of being converted per instruction
          const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
          if (tc::elect_one()) {
            auto kv_cols = [&](int j) { // keys of block j rounded up to the UMMA K / N granularity (16)
                const int rem = n_tok - j * BKV;
                return rem >= BKV ? BKV : ((rem + 15) & ~15);
            };
            // Issue cost matters (tools/mma_bench*.cu): building a shared-memory descriptor per tcgen05.mma costs ~130 clk per
            // instruction on the issuing thread, twice the 64 clk a 128x128x16 MMA runs.  All descriptors are therefore built
            // once; inside the unrolled K loops an MMA's operands are `base + compile-time constant`.
            const uint64_t qd0 = tc::make_smem_desc_sw128(tc::smem_u32(sQ), 16, 1024);
            const uint64_t pd0 = tc::make_smem_desc_sw128(tc::smem_u32(sP), 16, 1024);
            const uint64_t kd0 = tc::make_smem_desc_sw128(tc::smem_u32(sKV), 16, 1024);                 // K: K-major B operand
            // V is the MN-major B operand: 16 keys = 2 groups of 8 rows (SBO 1024 B); the two 64-wide head-dim boxes are LBO = 16 KB apart
            const uint64_t vd0 = tc::make_smem_desc_sw128(tc::smem_u32(sKV), TILE_BYTES / 2, 1024);
            constexpr uint64_t TILE_DESC = TILE_BYTES >> 4;  // descriptor start-address units are 16 bytes
            auto issue_S = [&](int t, int slot, int j) {    // caller has waited for Q_t, K(j) and for S_t to be free
                const uint64_t qd = qd0 + (uint64_t)t * TILE_DESC, kd = kd0 + (uint64_t)slot * TILE_DESC;
                const uint32_t idesc = tc::make_idesc_bf16(BQ, (uint32_t)kv_cols(j), 0, 0);
                const uint32_t d = tmem_base + TM_S + t * 128;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const uint64_t off = (uint64_t)(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4);
                    tc::umma_ss(d, qd + off, kd + off, idesc, k != 0 ? 1u : 0u);
                }
                tc::umma_commit(&s_full[t]);
            };
            constexpr uint32_t idesc_pv = tc::make_idesc_bf16(BQ, HD, 0, 1); // A = P K-major, B = V MN-major (head dim contiguous)
            auto issue_PV = [&](int t, int slot, int j) {   // caller has waited for P_t(j) and V(j)
                const uint64_t pd = pd0 + (uint64_t)t * TILE_DESC, vd = vd0 + (uint64_t)slot * TILE_DESC;
                const uint32_t o_tmem = tmem_base + TM_O + t * 128;
                const int ksteps = kv_cols(j) / 16;
                if (ksteps == BKV / 16) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        tc::umma_ss(o_tmem, pd + (uint64_t)(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4), vd + (uint64_t)((k * 2048) >> 4),
                                    idesc_pv, (j | k) != 0 ? 1u : 0u);
                } else {
                    for (int k = 0; k < ksteps; ++k)
                        tc::umma_ss(o_tmem, pd + (uint64_t)(((k >> 2) * (TILE_BYTES / 2) + (k & 3) * 32) >> 4), vd + (uint64_t)((k * 2048) >> 4),
                                    idesc_pv, (j | k) != 0 ? 1u : 0u);
                }
                tc::umma_commit(&pv_done[t]);
            };
            int c = 0;                                   // same K0, K1, V0, K2, V1, ... sequence as the producer
            int slot;
            auto next_tile = [&]() {
                slot = c % ATTN_RING;
                const uint32_t ph = (c / ATTN_RING) & 1;
                ++c;
                tc::mbar_wait(&kv_full[slot], ph);
            };
            int g = 0, it = 0;                           // iterations / items of MY tile so far (barrier phases)
            for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
                int q0, head, env; bool has_b;
                decode(w, q0, head, env, has_b);
                const bool mine = t == 0 || has_b;       // an item without a second tile: tile B's issuer only releases the ring slots
                next_tile();                             // K0
                if (mine) {
                    tc::mbar_wait(&q_full[t], it & 1);
                    if (g > 0) tc::mbar_wait(&s_free[t], (g - 1) & 1);   // the previous item's last S_t is in registers
                    tc::tc_fence_after();
                    issue_S(t, slot, 0);
                    if (n_kv == 1) tc::umma_commit(&q_empty[t]);
                    tc::umma_commit(&kv_empty[slot]);
                } else tc::mbar_arrive(&kv_empty[slot]);
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) {                      // S(j+1) as soon as the softmax warps hold S(j) in registers
                        next_tile();                         // K(j+1)
                        if (mine) {
                            tc::mbar_wait(&s_free[t], (g + j) & 1);
                            tc::tc_fence_after();
                            ATTN_TRACE(4, j, t == 0 ? 0 : 2);
                            issue_S(t, slot, j + 1);
                            if (j + 2 == n_kv) tc::umma_commit(&q_empty[t]);   // that was the item's last use of Q_t
                            tc::umma_commit(&kv_empty[slot]);
                        } else tc::mbar_arrive(&kv_empty[slot]);
                    }
                    next_tile();                             // V(j)
                    if (mine) {
                        if (j == 0) tc::mbar_wait(&o_free[t], (it & 1) ^ 1);   // the previous item's O_t has been written out
                        tc::mbar_wait(&p_ready[t], (g + j) & 1);
                        tc::tc_fence_after();
                        ATTN_TRACE(4, j, 3 + 2 * t);
                        issue_PV(t, slot, j);
                        ATTN_TRACE(4, j, 4 + 2 * t);
                        tc::umma_commit(&kv_empty[slot]);
                    } else tc::mbar_arrive(&kv_empty[slot]);
                }
                if (mine) { g += n_kv; ++it; }
            }
          }
        }
    } else {
        setmaxnreg_inc<192>();
        // ===================== softmax / correction / epilogue of tile t =====================
        const int t = (warp - 4) >> 2;
        const uint32_t tmem_base = *tmem_slot;
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
        const uint32_t s_addr = tmem_base + lane_addr + TM_S + t * 128;
        const uint32_t o_addr = tmem_base + lane_addr + TM_O + t * 128;
        // my row of P: 128-byte rows inside 1024-byte 8-row atoms, 16-byte chunk index XORed with (row & 7) (SWIZZLE_128B)
        uint8_t *p_row = sP + t * TILE_BYTES + row_in_tile * 128;
        const int sw = row_in_tile & 7;
        const float c = args.scale_log2;
        int g = 0;                    // iterations of this tile so far: every per-tile barrier completes one phase per iteration
        int n_lag = 0;                // items with two tiles so far (phase of `lag`)
        for (int w = blockIdx.x; w < args.n_items; w += gridDim.x) {
            int q0, head, env; bool has_b;
            decode(w, q0, head, env, has_b);
            if (t == 1 && !has_b) continue;
#ifndef SNB_ATTN_NO_LAG
            // De-phase the tiles.  Both softmaxes share the four XU (MUFU) pipes; started together they stay in phase for the whole item,
            // each exponential phase then runs at half the MUFU rate and the pipes idle during both tiles' load / max / store phases.
            // Tile B therefore starts its item half a block after tile A; nothing else couples the two, so the offset persists.
            if (t == 1) tc::mbar_wait(lag, n_lag & 1);
            if (has_b) ++n_lag;
#endif
            float m_used = -INFINITY; // running max the exponents are taken against (raw score units)
            float l = 0.0f;
            // one 128-key block.  RAGGED is a compile-time flag: ptxas if-converts the masking of keys that do not exist into 128 ISETP +
            // 128 SEL per row, a quarter of the loop's instructions -- only the copy that runs the item's LAST block carries them.
            auto block = [&](int j, auto ragged_tag) {
                constexpr bool RAGGED = decltype(ragged_tag)::value;
                const uint32_t ph = (g + j) & 1;
                ATTN_TRACE(t, j, 0);
                tc::mbar_wait(&s_full[t], ph);
                ATTN_TRACE(t, j, 1);
                __syncwarp();                 // reconverge before the .sync.aligned TMEM loads
                tc::tc_fence_after();
                // the whole S row: four 32-column loads in flight, ONE tcgen05.wait::ld
                uint32_t s0[32], s1[32], s2[32], s3[32];
                tc::tmem_ld_32x32(s_addr, s0);
                tc::tmem_ld_32x32(s_addr + 32, s1);
                tc::tmem_ld_32x32(s_addr + 64, s2);
                tc::tmem_ld_32x32(s_addr + 96, s3);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&s_free[t]);  // S_t may be overwritten by the next block's Q K^T
                ATTN_TRACE(t, j, 2);
                if constexpr (RAGGED) {            // ragged last block: keys that do not exist score -inf
                    const int valid = n_tok - j * BKV; // keys of this block that exist
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= valid) s0[i] = 0xff800000u;
                        if (32 + i >= valid) s1[i] = 0xff800000u;
                        if (64 + i >= valid) s2[i] = 0xff800000u;
                        if (96 + i >= valid) s3[i] = 0xff800000u;
                    }
                }
                float mx0 = __uint_as_float(s0[0]), mx1 = __uint_as_float(s1[0]), mx2 = __uint_as_float(s2[0]), mx3 = __uint_as_float(s3[0]);
#pragma unroll
                for (int i = 1; i < 31; i += 2) {
                    mx0 = fmax3(mx0, __uint_as_float(s0[i]), __uint_as_float(s0[i + 1])); mx1 = fmax3(mx1, __uint_as_float(s1[i]), __uint_as_float(s1[i + 1]));
                    mx2 = fmax3(mx2, __uint_as_float(s2[i]), __uint_as_float(s2[i + 1])); mx3 = fmax3(mx3, __uint_as_float(s3[i]), __uint_as_float(s3[i + 1]));
                }
                const float bmax = fmax3(fmax3(mx0, mx1, __uint_as_float(s0[31])), fmax3(mx2, mx3, __uint_as_float(s1[31])),
                                         fmaxf(__uint_as_float(s2[31]), __uint_as_float(s3[31])));
                if (j == 0) {
                    m_used = bmax;
                } else if (__any_sync(0xffffffffu, (bmax - m_used) * c > RESCALE_THRESHOLD)) {
                    // Refresh the running max.  The decision is WARP-uniform (tcgen05.ld/st are .sync.aligned and must be
                    // executed by all 32 lanes together); lanes whose own max did not grow rescale by exactly 1.
                    tc::mbar_wait(&pv_done[t], ph ^ 1);                   // O_t is stable once PV_t(j-1) has landed
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
                // exponentials first, into registers (packed bf16 pairs overwrite the scores they came from): pairs of keys on
                // FFMA2 / FADD2; one pair in POLY_EVERY takes the FMA-pipe polynomial instead of MUFU.EX2
                const uint64_t c2 = tc::f2_pack(c, c), nmc2 = tc::f2_pack(-mc, -mc);
                uint64_t sumA = tc::f2_pack(0.0f, 0.0f), sumB = sumA;
                auto exp_pack = [&](uint32_t (&s)[32]) {
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
                        s[i] = tc::pack_bf16(el, eh);
                    }
                };
                auto store = [&](const uint32_t (&s)[32], int box, int chunk0) {   // 32 keys = four 16-byte chunks of my P row
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4 *>(p_row + box * (TILE_BYTES / 2) + (((chunk0 + q) ^ sw) << 4)) =
                            make_uint4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
                };
                // Each 32-key quarter of the P row goes to shared memory as soon as its exponentials exist, so the LSU drains the
                // 32 KB tile (>= 256 clk of store bandwidth per tile) underneath the MUFU / FMA work of the next quarter instead of
                // after all of it (the trace of the exps-then-stores order: 471 clk of the block's 2930 on the stores alone).
                // PV_t(j-1) must have consumed the previous P first -- it was issued a whole softmax ago.
                // P(j) may only be stored once PV_t(j-1) has consumed P(j-1).  That MMA was issued a whole softmax ago but queues behind the
                // other tile's MMAs, and ncu's source view charged 8 % of all stall samples to this wait when it sat after the first
                // quarter; the packed exponentials stay in the registers of the scores they replace, so the wait (and the stores of the
                // quarters computed so far) can sit later at no register cost: SNB_ATTN_PV_WAIT_AFTER quarters of exponentials first.
                auto wait_p_free = [&]() {
                    if (g + j > 0) tc::mbar_wait(&pv_done[t], ph ^ 1);   // (for j == 0: the previous item's last PV, already awaited by its epilogue)
                    if (j == 0 && g > 0) {
                        // the previous item's O tile was staged in this P tile: its TMA store must have finished READING it
                        if (row_in_tile == 0) tc::tma_store_wait_read<0>();
                        tc::named_bar_sync(1 + t, 128);
                    }
                };
                exp_pack(s0);
                if (SNB_ATTN_PV_WAIT_AFTER == 1) { wait_p_free(); store(s0, 0, 0); }
#if !defined(SNB_ATTN_NO_LAG) && SNB_ATTN_LAG_AT == 0
                if (t == 0 && j == 0 && has_b) tc::mbar_arrive(lag);
#endif
                exp_pack(s1);
                if (SNB_ATTN_PV_WAIT_AFTER == 2) { wait_p_free(); store(s0, 0, 0); }
                if (SNB_ATTN_PV_WAIT_AFTER <= 2) store(s1, 0, 4);
#if !defined(SNB_ATTN_NO_LAG) && SNB_ATTN_LAG_AT == 1
                if (t == 0 && j == 0 && has_b) tc::mbar_arrive(lag);
#endif
                exp_pack(s2);
                if (SNB_ATTN_PV_WAIT_AFTER == 3) { wait_p_free(); store(s0, 0, 0); store(s1, 0, 4); }
                store(s2, 1, 0);
#if !defined(SNB_ATTN_NO_LAG) && SNB_ATTN_LAG_AT == 2
                if (t == 0 && j == 0 && has_b) tc::mbar_arrive(lag);
#endif
                exp_pack(s3);
                ATTN_TRACE(t, j, 3);
                store(s3, 1, 4);
#if !defined(SNB_ATTN_NO_LAG) && SNB_ATTN_LAG_AT == 3
                if (t == 0 && j == 0 && has_b) tc::mbar_arrive(lag);
#endif
                {
                    float a0, a1, b0, b1;
                    tc::f2_unpack(sumA, a0, a1);
                    tc::f2_unpack(sumB, b0, b1);
                    l += (a0 + a1) + (b0 + b1);
                }
                ATTN_TRACE(t, j, 4);
                tc::fence_proxy_async();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
                tc::mbar_arrive(&p_ready[t]);
                ATTN_TRACE(t, j, 5);
            };
            for (int j = 0; j < n_kv - 1; ++j) block(j, std::false_type{});
            if (n_tok % BKV) block(n_kv - 1, std::true_type{}); else block(n_kv - 1, std::false_type{});
            // final: O / l -> global
            tc::mbar_wait(&pv_done[t], (g + n_kv - 1) & 1);
            __syncwarp();
            tc::tc_fence_after();
            const float inv_l = 1.0f / l;
            // O / l goes out through the (now idle) P tile and ONE TMA store per 64-column box instead of per-thread 16-byte stores
            // to 32 different rows per instruction (trace of that version: 5 700 clk of the item's 46 000 in this epilogue).  Rows
            // beyond the sequence are clipped by the tensor map.
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t r[32];
                tc::tmem_ld_32x32(o_addr + ch * 32, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o;
                    o.x = tc::pack_bf16(__uint_as_float(r[8 * q + 0]) * inv_l, __uint_as_float(r[8 * q + 1]) * inv_l);
                    o.y = tc::pack_bf16(__uint_as_float(r[8 * q + 2]) * inv_l, __uint_as_float(r[8 * q + 3]) * inv_l);
                    o.z = tc::pack_bf16(__uint_as_float(r[8 * q + 4]) * inv_l, __uint_as_float(r[8 * q + 5]) * inv_l);
                    o.w = tc::pack_bf16(__uint_as_float(r[8 * q + 6]) * inv_l, __uint_as_float(r[8 * q + 7]) * inv_l);
                    *reinterpret_cast<uint4 *>(p_row + (ch >> 1) * (TILE_BYTES / 2) + ((((ch & 1) * 4 + q) ^ sw) << 4)) = o;
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&o_free[t]);      // O_t may be overwritten by the next item's first PV
            tc::fence_proxy_async();
            tc::named_bar_sync(1 + t, 128);   // the four warps of this tile: the whole O tile is in shared memory
            if (row_in_tile == 0) {
                tc::tma_store_3d(&tmOut, sP + t * TILE_BYTES, head * HD, q0 + t * BQ, env);
                tc::tma_store_3d(&tmOut, sP + t * TILE_BYTES + TILE_BYTES / 2, head * HD + 64, q0 + t * BQ, env);
                tc::tma_store_commit();
            }
            g += n_kv;
        }
        if (row_in_tile == 0) tc::tma_store_wait<0>();    // the last O tile has left shared memory before the CTA exits
    }
    __syncthreads();
#ifdef SNB_ATTN_TRACE
    if (trace_on && warp == 0) {
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_attn_trace[5 * 128 + 2] = clock64(); g_attn_trace[5 * 128 + 4] = gt;
    }
#endif
    if (warp == 2) tc::tmem_dealloc<TMEM_COLS>(__shfl_sync(0xffffffffu, *tmem_slot, 0));
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

int snb_attn_plan(AttnPlan *plan, const bf16 *qkv, bf16 *out, int n_env, int n_tok)
{
    SNB_REQUIRE(n_env > 0 && n_tok > 0, SNB_EINVAL, "attention: bad sizes");
    plan->n_env = n_env; plan->n_tok = n_tok; plan->out = out;
    int rc = snb_make_tmap_3d(&plan->tmQKV, qkv, (uint64_t)n_env, (uint64_t)n_tok, 1536, BKV);
    if (rc) return rc;
    return snb_make_tmap_3d(&plan->tmOut, out, (uint64_t)n_env, (uint64_t)n_tok, 512, BQ);   // O tiles: 128 rows x 64 columns per store
}

int snb_attn_launch(const AttnPlan *plan, cudaStream_t stream)
{
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM); });
    SNB_CUDA_TRY(attr_err);
    AttnArgs a;
    a.out = plan->out; a.n_tok = plan->n_tok;
    a.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
    a.n_qp = (plan->n_tok + 2 * BQ - 1) / (2 * BQ);
    a.n_items = a.n_qp * NHEAD * plan->n_env;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SNB_CUDA_TRY(cudaGetDevice(&dev));
        SNB_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = a.n_items < num_sms ? a.n_items : num_sms;   // persistent: one CTA per SM walks the item list
    attn_fwd_kernel<<<grid, ATTN_THREADS, ATTN_SMEM, stream>>>(plan->tmQKV, plan->tmOut, a);
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
