// csrc/jmid_kernels.cu -- the HBM-bound pieces of the JMID noise network around the tensor-core kernels:
// weight conversion, hyper-network (ConcatSquash gate / bias) tables, concat1 + positional encoding, LayerNorm,
// the 128 -> 2 output layer fused with the DDIM update, and the single-integrator cumsum.
// Reference: sicnav_diffusion/JMID/MID/models/common.py:37-72, models/diffusion.py:173-209, 507-531.
#include "jmid_internal.h"
#include "tc_utils.cuh"

namespace {

__global__ void f32_to_bf16_kernel(const float *__restrict__ src, bf16 *__restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// LayerNorm folded into the consuming nn.Linear (model load time): W'[n,k] = bf16(W[n,k] gamma[k]); colsum[n] = sum_k W'[n,k] (of the
// ROUNDED values: the GEMM multiplies by those); bias'[n] = bias[n] + sum_k beta[k] W[n,k].  One block per output row n.
__global__ void fold_ln_kernel(const float *__restrict__ W, const float *__restrict__ gamma, const float *__restrict__ beta,
                               const float *__restrict__ bias, bf16 *__restrict__ Wf, float *__restrict__ colsum, float *__restrict__ bias_f, int K)
{
    const int n = blockIdx.x;
    float cs = 0.0f, bb = 0.0f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float w = W[(size_t)n * K + k];
        const bf16 wf = __float2bfloat16_rn(w * gamma[k]);
        Wf[(size_t)n * K + k] = wf;
        cs += __bfloat162float(wf);
        bb = fmaf(beta[k], w, bb);
    }
    __shared__ float s_a[4], s_b[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cs += __shfl_xor_sync(0xffffffffu, cs, o); bb += __shfl_xor_sync(0xffffffffu, bb, o); }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = cs; s_b[threadIdx.x >> 5] = bb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        colsum[n] = (s_a[0] + s_a[1]) + (s_a[2] + s_a[3]);
        bias_f[n] = bias[n] + ((s_b[0] + s_b[1]) + (s_b[2] + s_b[3]));
    }
}

struct Hyper4 { HyperW l[4]; };

// HYPER_ROWS (ba) rows per block; threads stride over the 898 output columns; the ctx rows are staged in shared memory and every
// weight row a thread reads is used for all of the block's rows.  (One row per block re-read the 1.9 MB of hyper-network weights
// from L2 for each of a chunk's 5120 rows: 9.5 GB of L2 traffic, 2.5 ms per launch, ncu: SM throughput 12 %.)  The accumulation
// order per output is unchanged (k ascending, fmaf), so the tables keep their bits.
constexpr int HYPER_ROWS = 16;
__global__ void __launch_bounds__(256) hyper_ctx_kernel(const Hyper4 hw, const float *__restrict__ ctx, float *__restrict__ gc, float *__restrict__ bc,
                                                        int n_ba)
{
    __shared__ float s_ctx[HYPER_ROWS][256];
    const int ba0 = blockIdx.x * HYPER_ROWS;
    const int nr = min(HYPER_ROWS, n_ba - ba0);
    for (int k = threadIdx.x; k < HYPER_ROWS * 256; k += blockDim.x) {
        const int r = k >> 8;
        s_ctx[r][k & 255] = r < nr ? ctx[(size_t)(ba0 + r) * 256 + (k & 255)] : 0.0f;
    }
    __syncthreads();
    for (int col = threadIdx.x; col < HYPER_TOTAL; col += blockDim.x) {
        int li = 0, n = col;
        while (n >= hw.l[li].dout) { n -= hw.l[li].dout; ++li; }
        const float *wg = hw.l[li].gate_w + (size_t)n * 259 + 3;
        const float *wb = hw.l[li].bias_w + (size_t)n * 259 + 3;
        float g[HYPER_ROWS], b[HYPER_ROWS];
        const float g0 = hw.l[li].gate_b[n];
#pragma unroll
        for (int r = 0; r < HYPER_ROWS; ++r) { g[r] = g0; b[r] = 0.0f; }
        for (int k = 0; k < 256; ++k) {
            const float vg = wg[k], vb = wb[k];
#pragma unroll
            for (int r = 0; r < HYPER_ROWS; ++r) { g[r] = fmaf(vg, s_ctx[r][k], g[r]); b[r] = fmaf(vb, s_ctx[r][k], b[r]); }
        }
#pragma unroll
        for (int r = 0; r < HYPER_ROWS; ++r)
            if (r < nr) {
                gc[(size_t)(ba0 + r) * HYPER_LD + col] = g[r];
                bc[(size_t)(ba0 + r) * HYPER_LD + col] = b[r];
            }
    }
}

// block = 32 columns x 8 row lanes; a thread fetches the six time-embedding weights of its column once (rows of stride 259 floats:
// scattered) and walks HYPER_IT_ROWS / 8 rows with them; a warp touches 128 contiguous bytes of one table row per access.  (One thread
// per element re-fetched the six scattered weights for every element: 93 us per call for 74 MB of table traffic.)
constexpr int HYPER_IT_ROWS = 64;
__global__ void __launch_bounds__(256) hyper_iter_kernel(const Hyper4 hw, const float *__restrict__ gc, const float *__restrict__ bc,
                                                         float *__restrict__ gate, float *__restrict__ hb, int n_ba, float beta, float sb, float cb)
{
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    if (col >= HYPER_TOTAL) return;
    int li = 0, n = col;
    while (n >= hw.l[li].dout) { n -= hw.l[li].dout; ++li; }
    const float *wg = hw.l[li].gate_w + (size_t)n * 259;
    const float *wb = hw.l[li].bias_w + (size_t)n * 259;
    const float wg0 = wg[0], wg1 = wg[1], wg2 = wg[2], wb0 = wb[0], wb1 = wb[1], wb2 = wb[2];
    const int r0 = blockIdx.y * HYPER_IT_ROWS + (threadIdx.x >> 5);
    const int r1 = min(n_ba, (int)(blockIdx.y + 1) * HYPER_IT_ROWS);
    for (int r = r0; r < r1; r += 8) {
        const size_t i = (size_t)r * HYPER_LD + col;
        const float g = gc[i] + wg0 * beta + wg1 * sb + wg2 * cb;
        gate[i] = 1.0f / (1.0f + __expf(-g));
        hb[i] = bc[i] + wb0 * beta + wb1 * sb + wb2 * cb;
    }
}

// concat1 (2 -> 512) + gate/bias + positional encoding.  One CTA of 128 threads per (env, agent): the thread's 4 columns of the
// gate / hyper-bias row, of W1, b1 and (T <= 8) of the positional encoding stay in registers while it walks the S*T tokens that
// share them; per token that leaves two broadcast loads and one coalesced 8-byte store (the first version re-read 96 B of
// tables per 8 B written and ran at a third of the HBM write rate).
__global__ void __launch_bounds__(128) embed_kernel(const float *__restrict__ x, const float *__restrict__ w1, const float *__restrict__ b1,
                                                    const float *__restrict__ gate, const float *__restrict__ hb, const float *__restrict__ pe,
                                                    bf16 *__restrict__ h, int n_ba, int tok_per_env, int T, int A)
{
    const int ba = blockIdx.x;
    if (ba >= n_ba) return;
    const int b = ba / A, a = ba - b * A;
    const int S = tok_per_env / (A * T);
    const int c4 = threadIdx.x * 4;
    const float4 g = *reinterpret_cast<const float4 *>(gate + (size_t)ba * HYPER_LD + c4);
    const float4 hbv = *reinterpret_cast<const float4 *>(hb + (size_t)ba * HYPER_LD + c4);
    const float4 bb = *reinterpret_cast<const float4 *>(b1 + c4);
    const float4 wa = *reinterpret_cast<const float4 *>(w1 + 2 * c4);     // w1[c4][0], w1[c4][1], w1[c4+1][0], w1[c4+1][1]
    const float4 wb = *reinterpret_cast<const float4 *>(w1 + 2 * c4 + 4);
    // (W1 x + b1) * g + hb + pe  =  x0 * (w?0 g) + x1 * (w?1 g) + (b g + hb) + pe
    const float k00 = wa.x * g.x, k01 = wa.y * g.x, k0 = bb.x * g.x + hbv.x;
    const float k10 = wa.z * g.y, k11 = wa.w * g.y, k1 = bb.y * g.y + hbv.y;
    const float k20 = wb.x * g.z, k21 = wb.y * g.z, k2 = bb.z * g.z + hbv.z;
    const float k30 = wb.z * g.w, k31 = wb.w * g.w, k3 = bb.w * g.w + hbv.w;
    for (int tau = 0; tau < T; ++tau) {
        const float4 p = *reinterpret_cast<const float4 *>(pe + (size_t)tau * 512 + c4);
        const float q0 = k0 + p.x, q1 = k1 + p.y, q2 = k2 + p.z, q3 = k3 + p.w;
#pragma unroll 4
        for (int sidx = 0; sidx < S; ++sidx) {
            const size_t tok = (size_t)b * tok_per_env + (size_t)(sidx * A + a) * T + tau;
            const float2 xv = *reinterpret_cast<const float2 *>(x + 2 * tok);
            uint2 o;
            o.x = tc::pack_bf16(fmaf(xv.y, k01, fmaf(xv.x, k00, q0)), fmaf(xv.y, k11, fmaf(xv.x, k10, q1)));
            o.y = tc::pack_bf16(fmaf(xv.y, k21, fmaf(xv.x, k20, q2)), fmaf(xv.y, k31, fmaf(xv.x, k30, q3)));
            *reinterpret_cast<uint2 *>(h + tok * 512 + c4) = o;
        }
    }
}

// out = LayerNorm(pre + resid) over 512 columns, one warp per row (eps 1e-5, biased variance like torch); `pre` (the
// sub-layer output written by the GEMM epilogue) and `resid` are bf16, statistics and normalisation are fp32
__global__ void layernorm_kernel(const bf16 *__restrict__ pre, const bf16 *__restrict__ resid, const float *__restrict__ g,
                                 const float *__restrict__ b, bf16 *__restrict__ out, int rows)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint4 *pp = reinterpret_cast<const uint4 *>(pre + (size_t)row * 512);
    const uint4 *rp = reinterpret_cast<const uint4 *>(resid + (size_t)row * 512);
    float v[16];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint4 a = pp[lane + 32 * i], r = rp[lane + 32 * i];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162 *>(&aw[q]), y = *reinterpret_cast<const __nv_bfloat162 *>(&rw[q]);
            v[8 * i + 2 * q] = __bfloat162float(x.x) + __bfloat162float(y.x);
            v[8 * i + 2 * q + 1] = __bfloat162float(x.y) + __bfloat162float(y.y);
            sum += v[8 * i + 2 * q] + v[8 * i + 2 * q + 1];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / 512.0f);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; sq += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / 512.0f) + 1e-5f);
    uint4 *op = reinterpret_cast<uint4 *>(out + (size_t)row * 512);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4 *>(g + c0)), g1 = __ldg(reinterpret_cast<const float4 *>(g + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(b + c0)), b1 = __ldg(reinterpret_cast<const float4 *>(b + c0 + 4));
        uint4 o;
        o.x = tc::pack_bf16((v[8 * i + 0] - mean) * rstd * g0.x + b0.x, (v[8 * i + 1] - mean) * rstd * g0.y + b0.y);
        o.y = tc::pack_bf16((v[8 * i + 2] - mean) * rstd * g0.z + b0.z, (v[8 * i + 3] - mean) * rstd * g0.w + b0.w);
        o.z = tc::pack_bf16((v[8 * i + 4] - mean) * rstd * g1.x + b1.x, (v[8 * i + 5] - mean) * rstd * g1.y + b1.y);
        o.w = tc::pack_bf16((v[8 * i + 6] - mean) * rstd * g1.z + b1.z, (v[8 * i + 7] - mean) * rstd * g1.w + b1.w);
        op[lane + 32 * i] = o;
    }
}

// `linear` ConcatSquash 128 -> 2 on the concat4 output, then the DDIM update (diffusion.py:524-528).
// 8 lanes per token (16 columns each), 4 tokens per warp; a thread keeps the 32 weights of its 16 columns in registers and walks
// TAIL_TOK_ITERS tokens with them (it used to issue 64 scalar weight loads per token next to the two 16-byte data loads).  Same
// accumulation order per lane and the same shuffle tree as before, hence the same bits.
constexpr int TAIL_TOK_ITERS = 8;
__global__ void __launch_bounds__(256) tail_ddim_kernel(const bf16 *__restrict__ t4, const float *__restrict__ wl, const float *__restrict__ bl,
                                                        const float *__restrict__ gate, const float *__restrict__ hb, int tab_ld,
                                                        const float *__restrict__ x_t, float *__restrict__ x_next, float *__restrict__ eps_out,
                                                        int n_tok_total, int tok_per_env, int T, int A, float c1, float c2, float c3, float c4)
{
    const int sub = threadIdx.x & 7;
    float w0[16], w1[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(wl + sub * 16) + q);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(wl + 128 + sub * 16) + q);
        w0[4 * q] = a.x; w0[4 * q + 1] = a.y; w0[4 * q + 2] = a.z; w0[4 * q + 3] = a.w;
        w1[4 * q] = b.x; w1[4 * q + 1] = b.y; w1[4 * q + 2] = b.z; w1[4 * q + 3] = b.w;
    }
    const float bl0 = bl[0], bl1 = bl[1];
    const int tok0 = blockIdx.x * (32 * TAIL_TOK_ITERS) + (threadIdx.x >> 3);
#pragma unroll 2
    for (int it = 0; it < TAIL_TOK_ITERS; ++it) {
        const int tok = tok0 + it * 32;
        const bool ok = tok < n_tok_total;
        float a0 = 0.0f, a1 = 0.0f;
        if (ok) {
            const uint4 *p = reinterpret_cast<const uint4 *>(t4 + (size_t)tok * 128 + sub * 16);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint4 u = p[q];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
                    const int k = q * 8 + 2 * i;
                    const float f0 = __bfloat162float(v.x), f1 = __bfloat162float(v.y);
                    a0 = fmaf(f0, w0[k], a0); a0 = fmaf(f1, w0[k + 1], a0);
                    a1 = fmaf(f0, w1[k], a1); a1 = fmaf(f1, w1[k + 1], a1);
                }
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
        if (ok && sub == 0) {
            const int b = tok / tok_per_env;
            const int r = (tok - b * tok_per_env) / T;
            const int ba = b * A + (r % A);
            const float *gp = gate + (size_t)ba * tab_ld, *hp = hb + (size_t)ba * tab_ld;
            const float e0 = (a0 + bl0) * gp[0] + hp[0];
            const float e1 = (a1 + bl1) * gp[1] + hp[1];
            if (eps_out) { eps_out[2 * (size_t)tok] = e0; eps_out[2 * (size_t)tok + 1] = e1; }
            if (x_next) {
                const float x0 = x_t[2 * (size_t)tok], x1 = x_t[2 * (size_t)tok + 1];
                const float p0 = (x0 - e0 * c1) / c2, p1 = (x1 - e1 * c1) / c2;   // x0_t
                x_next[2 * (size_t)tok] = c3 * p0 + c4 * e0;
                x_next[2 * (size_t)tok + 1] = c3 * p1 + c4 * e1;
            }
        }
    }
}

// positions = cumsum_t(v) * dt + p0[a]   (single_integrator.py:321); one thread per (b, s, a, xy)
__global__ void integrate_kernel(const float *__restrict__ vel, const float *__restrict__ p0, float *__restrict__ pos, int B, int S,
                                 int A, int T, float dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * S * A * 2) return;
    const int c = i & 1;
    const int a = (i >> 1) % A;
    const int b = (i >> 1) / (A * S);
    const size_t base = (size_t)(i >> 1) * T * 2 + c;
    const float start = p0[((size_t)b * A + a) * 2 + c];
    float acc = 0.0f;
    for (int t = 0; t < T; ++t) {
        acc += vel[base + 2 * t];
        pos[base + 2 * t] = acc * dt + start;
    }
}

} // namespace

int snb_k_fold_ln(const float *W, const float *gamma, const float *beta, const float *bias, bf16 *Wf, float *colsum, float *bias_f,
                  int N, int K, cudaStream_t s)
{
    fold_ln_kernel<<<N, 128, 0, s>>>(W, gamma, beta, bias, Wf, colsum, bias_f, K);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_f32_to_bf16(const float *src, bf16 *dst, size_t n, cudaStream_t s)
{
    f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, n);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

static Hyper4 make_h4(const HyperW *l4)
{
    Hyper4 h;
    for (int i = 0; i < 4; ++i) h.l[i] = l4[i];
    return h;
}

int snb_k_hyper_ctx(const HyperW *layers4, const float *ctx, float *gc, float *bc, int n_ba, cudaStream_t s)
{
    hyper_ctx_kernel<<<(n_ba + HYPER_ROWS - 1) / HYPER_ROWS, 256, 0, s>>>(make_h4(layers4), ctx, gc, bc, n_ba);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_hyper_iter(const HyperW *layers4, const float *gc, const float *bc, float *gate, float *hb, int n_ba, float beta, cudaStream_t s)
{
    const dim3 grid((HYPER_TOTAL + 31) / 32, (n_ba + HYPER_IT_ROWS - 1) / HYPER_IT_ROWS);
    hyper_iter_kernel<<<grid, 256, 0, s>>>(make_h4(layers4), gc, bc, gate, hb, n_ba, beta, sinf(beta), cosf(beta));
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_embed(const float *x, const float *w1, const float *b1, const float *gate, const float *hb, const float *pe, bf16 *h,
                int n_tok_total, int tok_per_env, int T, int A, cudaStream_t s)
{
    embed_kernel<<<(n_tok_total / tok_per_env) * A, 128, 0, s>>>(x, w1, b1, gate, hb, pe, h, (n_tok_total / tok_per_env) * A, tok_per_env, T, A);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_layernorm(const bf16 *in, const bf16 *resid, const float *g, const float *b, bf16 *out, int rows, cudaStream_t s)
{
    layernorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(in, resid, g, b, out, rows);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_tail_ddim(const bf16 *t4, const float *wl, const float *bl, const float *gate, const float *hb, int tab_ld,
                    const float *x_t, float *x_next, float *eps_out, int n_tok_total, int tok_per_env, int T, int A,
                    float c1, float c2, float c3, float c4, cudaStream_t s)
{
    const int tok_per_block = 32 * TAIL_TOK_ITERS;
    tail_ddim_kernel<<<(unsigned)((n_tok_total + tok_per_block - 1) / tok_per_block), 256, 0, s>>>(t4, wl, bl, gate, hb, tab_ld, x_t, x_next, eps_out, n_tok_total,
                                                                       tok_per_env, T, A, c1, c2, c3, c4);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_integrate(const float *vel, const float *p0, float *pos, int B, int S, int A, int T, float dt, cudaStream_t s)
{
    const int n = B * S * A * 2;
    integrate_kernel<<<(n + 255) / 256, 256, 0, s>>>(vel, p0, pos, B, S, A, T, dt);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
