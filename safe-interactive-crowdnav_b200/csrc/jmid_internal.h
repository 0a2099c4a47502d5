// csrc/jmid_internal.h -- internal interfaces between the denoiser translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "snb_common.h"

typedef __nv_bfloat16 bf16;

// ---- TMA tensor maps (jmid_gemm.cu) ----
// 2-D row-major bf16 matrix [rows, cols] (cols contiguous), box = {64 cols, box_rows}, 128-byte swizzle.
int snb_make_tmap_2d(CUtensorMap *out, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
// 3-D bf16 tensor [d2, d1, d0] (d0 contiguous), box = {64, box_rows, 1}, 128-byte swizzle.
int snb_make_tmap_3d(CUtensorMap *out, const void *ptr, uint64_t d2, uint64_t d1, uint64_t d0, uint32_t box_rows);

// ---- GEMM: C[M,N] = A[M,K] * W[N,K]^T (+ epilogue), bf16 operands, fp32 accumulation in TMEM ----
enum GemmEpiKind {
    EPI_BIAS_BF16 = 0,      // out_bf16 = acc + bias[n]
    EPI_BIAS_RELU_BF16 = 1, // out_bf16 = relu(acc + bias[n])
    EPI_BIAS_F32 = 2,       // out_f32  = acc + bias[n]   (pre-LayerNorm; the residual is added by the LayerNorm kernel)
    EPI_CSL_BF16 = 3        // out_bf16 = (acc + bias[n]) * gate[ba(m),n] + hbias[ba(m),n]   (ConcatSquashLinear)
};

struct GemmEpi {
    const float *bias;      // [N]
    const float *gate;      // [n_ba, tab_ld] already sigmoid()ed, pointing at this layer's first column
    const float *hbias;     // [n_ba, tab_ld]
    int tab_ld;
    int tok_per_env, T, A;  // row m -> b = m / tok_per_env, r = (m % tok_per_env) / T, a = r % A, ba = b*A + a
    // LayerNorm fold (jmid_gemm.cu, CTA-pair kernel): all NULL / 0 = plain epilogue
    int fold;                 // 1: the A operand is a pre-norm z; `bias` is beta W^T + b, `colsum` = colsum(W diag(gamma)), `stats_in` [M,4] of z
    int res;                  // 1: + resid (bf16 [M,N]);  2: + LayerNorm(resid) from res_stats / res_gamma / res_beta;  both write stats_out
    const float *colsum;
    const float2 *stats_in;
    const bf16 *resid;
    const float2 *res_stats;
    const float *res_gamma, *res_beta;
    float2 *stats_out;        // [M,4] (sum, sum of squares) per 128-column slice of the written row
};

struct GemmPlan {
    CUtensorMap tmA, tmB, tmB2, tmC; // tmB2: W with box BN/2 rows (CTA-pair kernel); tmC: output [M,N] (bf16 or fp32), written with TMA stores of 32 rows x 128 B
    int M, N, K, BN, out_f32;
};
int snb_gemm_plan(GemmPlan *plan, const bf16 *A, const bf16 *W, void *out, int out_f32, int M, int N, int K);
int snb_gemm_launch(const GemmPlan *plan, int epi_kind, const GemmEpi *epi, int num_sms, cudaStream_t stream);

// ---- attention (jmid_attn.cu) ----
// qkv [n_env, n_tok, 1536] bf16 (Q | K | V, 4 heads x 128 each) -> out [n_env * n_tok, 512] bf16; one unmasked
// sequence of n_tok tokens per environment.
struct AttnPlan {
    CUtensorMap tmQKV, tmOut;   // out [n_env, n_tok, 512]: the O tiles leave through TMA stores (rows beyond the sequence are clipped)
    bf16 *out;
    int n_env, n_tok;
};
int snb_attn_plan(AttnPlan *plan, const bf16 *qkv, bf16 *out, int n_env, int n_tok);
int snb_attn_launch(const AttnPlan *plan, cudaStream_t stream);
// second-generation kernel (jmid_attn2.cu): 64-key blocks, P in TMEM, ten-slot K / V ring
struct Attn2Plan {
    CUtensorMap tmQ, tmKV;   // boxes of 128 (queries) and 64 (keys) rows x 64 columns
    int n_env, n_tok;
};
int snb_attn2_plan(Attn2Plan *plan, const bf16 *qkv, int n_env, int n_tok);
int snb_attn2_launch(const Attn2Plan *plan, bf16 *out, cudaStream_t stream);
// iMID: independent sequences of T (<= 32) tokens; qkv [n_seq * T, 1536] -> out [n_seq * T, 512]
int snb_attn_small_launch(const bf16 *qkv, bf16 *out, int n_seq, int T, cudaStream_t stream);

// ---- elementwise kernels (jmid_kernels.cu) ----
int snb_k_f32_to_bf16(const float *src, bf16 *dst, size_t n, cudaStream_t s);
// LayerNorm(gamma, beta) folded into the consuming Linear(W [N,K], bias): Wf = bf16(W diag(gamma)), colsum(Wf), bias + beta W^T
int snb_k_fold_ln(const float *W, const float *gamma, const float *beta, const float *bias, bf16 *Wf, float *colsum, float *bias_f,
                  int N, int K, cudaStream_t s);
// ctx-dependent, iteration-invariant part of the 4 hyper networks: gc[ba, 898] = Wg[:,3:] ctx[ba] + bg, bc = Wb[:,3:] ctx[ba]
struct HyperW { const float *gate_w, *gate_b, *bias_w; int dout; }; // gate_w / bias_w are [dout, 259]
int snb_k_hyper_ctx(const HyperW *layers4, const float *ctx, float *gc, float *bc, int n_ba, cudaStream_t s);
// per iteration: gate = sigmoid(gc + Wg[:, :3] temb), hb = bc + Wb[:, :3] temb, temb = (beta, sin beta, cos beta)
int snb_k_hyper_iter(const HyperW *layers4, const float *gc, const float *bc, float *gate, float *hb, int n_ba, float beta, cudaStream_t s);
// concat1 + positional encoding: h[m, 512] = (W1 x[m] + b1) * gate[ba, 0:512] + hb[ba, 0:512] + pe[tau]
int snb_k_embed(const float *x, const float *w1, const float *b1, const float *gate, const float *hb, const float *pe, bf16 *h,
                int n_tok_total, int tok_per_env, int T, int A, cudaStream_t s);
// out_bf16 = LayerNorm(512)(in_bf16 + resid_bf16), fp32 statistics
int snb_k_layernorm(const bf16 *in, const bf16 *resid, const float *g, const float *b, bf16 *out, int rows, cudaStream_t s);
// final ConcatSquash 128 -> 2 and the DDIM update of x_t (diffusion.py:524-528); eps_out optional
int snb_k_tail_ddim(const bf16 *t4, const float *wl, const float *bl, const float *gate, const float *hb, int tab_ld,
                    const float *x_t, float *x_next, float *eps_out, int n_tok_total, int tok_per_env, int T, int A,
                    float c_sqrt_1mab, float c_sqrt_ab, float c_sqrt_abn, float c_sqrt_1mabn, cudaStream_t s);
int snb_k_integrate(const float *vel, const float *p0, float *pos, int B, int S, int A, int T, float dt, cudaStream_t s);

// ---- fp32-class path (jmid_fp32x.cu) ----
int snb_x_split3(const float *src, bf16 *dst, size_t rows, int K, int weight_order, int relu, cudaStream_t s);
int snb_x_embed(const float *x, const float *w1, const float *b1, const float *gate, const float *hb, const float *pe, float *h,
                int n_tok_total, int tok_per_env, int T, int A, cudaStream_t s);
int snb_x_layernorm(const float *pre, const float *resid, const float *g, const float *b, float *out, int rows, cudaStream_t s);
int snb_x_csl_apply(float *v, const float *gate, const float *hb, size_t rows, int N, int tok_per_env, int T, int A, cudaStream_t s);
int snb_x_attention(const float *qkv, float *out, int n_seq, int seq_len, cudaStream_t s);
int snb_x_tail_ddim(const float *t4, const float *wl, const float *bl, const float *gate, const float *hb, const float *x_t, float *x_next,
                    float *eps_out, int n_tok_total, int tok_per_env, int T, int A, float c1, float c2, float c3, float c4, cudaStream_t s);

#define HYPER_TOTAL 898 // 512 + 256 + 128 + 2 columns of the four hyper networks
#define HYPER_LD 900    // row stride of the gate / bias tables (16-byte aligned rows)
