// csrc/jmid_api.cu -- C-ABI entry points of the JMID / iMID denoiser (include/snb.h): model construction,
// the batched DDIM sampling loop (DiffusionTraj.sample_sicnav_inference, models/diffusion.py:478-541), one-shot
// eps, integration and the host-buffer plugin call.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "jmid_internal.h"

namespace {

constexpr int D = 512, DFF = 1024, NL = 3;

// SNB_ATTN_V2=1 selects the experimental second kernel (jmid_attn2.cu: 64-key blocks, P in TMEM) for A/B runs; measured on B200 it is
// NOT faster (1005 vs 1038 TFLOP/s): its N = 64 S = Q K^T MMAs re-read the 4 KB Q slice from shared memory for half the math and run at
// 48 clk instead of 32 (the 128 B / clk operand port), so the tensor side becomes the bound.  Default: jmid_attn.cu.
bool attn_v1()
{
    static const bool v2 = [] { const char *e = getenv("SNB_ATTN_V2"); return e && e[0] == '1'; }();
    return !v2;
}

struct LayerDev {
    bf16 *wqkv, *wo, *w1, *w2;
    float *bqkv, *bo, *b1, *b2, *n1w, *n1b, *n2w, *n2b;
    // LayerNorm fold: linear1 with norm1 of this layer folded in; in_proj with norm2 of the PREVIOUS layer (layers 1, 2)
    bf16 *w1_f, *wqkv_f;
    float *cs_1, *b1_f, *cs_qkv, *bqkv_f;
};

struct Plans {
    int n_env = 0, M = 0, A = 0, N = 0; // A agents per env in this call (<= the handle's A), N = A*S*T tokens per env
    GemmPlan qkv[NL], out[NL], ff1[NL], ff2[NL], c3, c4;
    GemmPlan qkv_f[NL], ff1_f[NL], ff2_f[NL], c3_f;   // LayerNorm-fold dataflow: z1 lives in `pre`, z2 in `y`
    AttnPlan attn;
    Attn2Plan attn2;
};

} // namespace

struct SnbJmid {
    int A, S, T, joint, max_envs, chunk_envs, N /*tokens per env*/, num_sms;
    std::vector<void *> allocs;
    // weights
    LayerDev L[NL];
    bf16 *wc3, *wc4, *wc3_f = nullptr;
    float *cs_c3 = nullptr, *bc3_f = nullptr;
    float2 *stats1 = nullptr, *stats2 = nullptr;   // [M,4] (sum, sum of squares) per 128-column slice of z1 / z2
    int ln_fold = 1;
    float *c1_w, *c1_b, *c3_b, *c4_b, *lin_w, *lin_b, *pe;
    HyperW hyper[4];
    float betas[101], alpha_bars[101];
    // activations for one chunk
    bf16 *h, *y, *qkv, *att, *ff, *t3, *t4;
    bf16 *pre;
    float *xa, *xb, *gc, *bc, *gate, *hb;
    std::map<std::pair<int, int>, Plans> plans; // keyed by (envs in chunk, agents per env)
    // CUDA graphs of one chunk's whole DDIM loop, keyed by (envs in chunk, agents per env, n_steps); captured on second use
    float *ctx_stage = nullptr;
    std::map<int64_t, cudaGraphExec_t> graphs;
    std::map<int64_t, int> graph_uses;
    int use_graphs = 1;
    cudaStream_t own_stream = nullptr;
    // precision "fp32x" (csrc/jmid_fp32x.cu): fp32 activations, weights split into 3 bf16 pieces laid out along K (K' = 6K)
    int precision = SNB_PREC_BF16, x_chunk_envs = 0;
    bf16 *x_wqkv[NL] = {}, *x_wo[NL] = {}, *x_w1[NL] = {}, *x_w2[NL] = {}, *x_wc3 = nullptr, *x_wc4 = nullptr, *x_a6 = nullptr;
    float *x_h = nullptr, *x_y = nullptr, *x_qkv = nullptr, *x_att = nullptr, *x_ff = nullptr, *x_pre = nullptr, *x_t3 = nullptr, *x_t4 = nullptr;
    const float *src_w[NL][4] = {}, *src_c3 = nullptr, *src_c4 = nullptr;   // caller's fp32 weights are copied at create (kept for the split)
    std::map<std::pair<int, int>, Plans> x_plans;
    // host-call staging
    float *d_ctx = nullptr, *d_xT = nullptr, *d_p0 = nullptr, *d_vel = nullptr, *d_pos = nullptr;
};

namespace {

template <class Tp>
int dev_alloc(SnbJmid *h, Tp **p, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(Tp));
    if (e != cudaSuccess) { snb_set_error("snb_jmid: cudaMalloc(%zu B) failed: %s", n * sizeof(Tp), cudaGetErrorString(e)); return SNB_ENOMEM; }
    h->allocs.push_back(q);
    *p = static_cast<Tp *>(q);
    return SNB_OK;
}

int dup_f32(SnbJmid *h, float **dst, const float *src, size_t n, cudaStream_t s)
{
    SNB_REQUIRE(src != nullptr, SNB_EINVAL, "snb_jmid_create: NULL weight pointer");
    int rc = dev_alloc(h, dst, n);
    if (rc) return rc;
    SNB_CUDA_TRY(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return SNB_OK;
}

int dup_bf16(SnbJmid *h, bf16 **dst, const float *src, size_t n, cudaStream_t s)
{
    SNB_REQUIRE(src != nullptr, SNB_EINVAL, "snb_jmid_create: NULL weight pointer");
    int rc = dev_alloc(h, dst, n);
    if (rc) return rc;
    return snb_k_f32_to_bf16(src, *dst, n, s);
}

int get_plans(SnbJmid *h, int n_env, int A, Plans **out)
{
    const std::pair<int, int> key(n_env, A);
    auto it = h->plans.find(key);
    if (it != h->plans.end()) { *out = &it->second; return SNB_OK; }
    Plans p;
    p.n_env = n_env;
    p.A = A;
    p.N = A * h->S * h->T;
    p.M = n_env * p.N;
    int rc = 0;
    for (int l = 0; l < NL && !rc; ++l) {
        rc = snb_gemm_plan(&p.qkv[l], h->h, h->L[l].wqkv, h->qkv, 0, p.M, 3 * D, D);
        if (!rc) rc = snb_gemm_plan(&p.out[l], h->att, h->L[l].wo, h->pre, 0, p.M, D, D);
        if (!rc) rc = snb_gemm_plan(&p.ff1[l], h->y, h->L[l].w1, h->ff, 0, p.M, DFF, D);
        if (!rc) rc = snb_gemm_plan(&p.ff2[l], h->ff, h->L[l].w2, h->pre, 0, p.M, D, DFF);
    }
    if (h->ln_fold) {
        for (int l = 0; l < NL && !rc; ++l) {
            if (l > 0) rc = snb_gemm_plan(&p.qkv_f[l], h->y, h->L[l].wqkv_f, h->qkv, 0, p.M, 3 * D, D);
            if (!rc) rc = snb_gemm_plan(&p.ff1_f[l], h->pre, h->L[l].w1_f, h->ff, 0, p.M, DFF, D);
            if (!rc) rc = snb_gemm_plan(&p.ff2_f[l], h->ff, h->L[l].w2, h->y, 0, p.M, D, DFF);
        }
        if (!rc) rc = snb_gemm_plan(&p.c3_f, h->y, h->wc3_f, h->t3, 0, p.M, 256, D);
    }
    if (!rc) rc = snb_gemm_plan(&p.c3, h->h, h->wc3, h->t3, 0, p.M, 256, D);
    if (!rc) rc = snb_gemm_plan(&p.c4, h->t3, h->wc4, h->t4, 0, p.M, 128, 256);
    if (!rc && h->joint) rc = snb_attn_plan(&p.attn, h->qkv, h->att, n_env, p.N);
    if (!rc && h->joint) rc = snb_attn2_plan(&p.attn2, h->qkv, n_env, p.N);
    if (rc) return rc;
    h->plans[key] = p;
    *out = &h->plans[key];
    return SNB_OK;
}

// one noise-network forward for the chunk currently staged in h->gc / h->bc / x_in (diffusion.py:173-209)
int net_forward(SnbJmid *h, Plans *P, const float *x_in, float *x_next, float *eps_out, int t, int t_next, cudaStream_t s)
{
    const int M = P->M, n_ba = P->n_env * P->A;
    int rc = snb_k_hyper_iter(h->hyper, h->gc, h->bc, h->gate, h->hb, n_ba, h->betas[t], s);
    if (rc) return rc;
    rc = snb_k_embed(x_in, h->c1_w, h->c1_b, h->gate, h->hb, h->pe, h->h, M, P->N, h->T, P->A, s);
    if (rc) return rc;
    GemmEpi e;
    if (h->ln_fold) {
        // LayerNorm folded into the GEMMs (jmid_gemm.cu): z1 = h + attn -> `pre` (+ stats1), z2 = LN1(z1) + ff -> `y` (+ stats2); no
        // LayerNorm kernel, no normalised activation is ever written to HBM.
        for (int l = 0; l < NL; ++l) {
            memset(&e, 0, sizeof(e));
            if (l == 0) { e.bias = h->L[0].bqkv; rc = snb_gemm_launch(&P->qkv[0], EPI_BIAS_BF16, &e, h->num_sms, s); }
            else {
                e.bias = h->L[l].bqkv_f; e.fold = 1; e.colsum = h->L[l].cs_qkv; e.stats_in = h->stats2;
                rc = snb_gemm_launch(&P->qkv_f[l], EPI_BIAS_BF16, &e, h->num_sms, s);
            }
            if (rc) return rc;
            if (h->joint) rc = attn_v1() ? snb_attn_launch(&P->attn, s) : snb_attn2_launch(&P->attn2, h->att, s);
            else rc = snb_attn_small_launch(h->qkv, h->att, M / h->T, h->T, s);
            if (rc) return rc;
            memset(&e, 0, sizeof(e));                       // out-proj: z1 = attn W_o^T + b_o + (layer input)
            e.bias = h->L[l].bo; e.stats_out = h->stats1;
            if (l == 0) { e.res = 1; e.resid = h->h; }
            else { e.res = 2; e.resid = h->y; e.res_stats = h->stats2; e.res_gamma = h->L[l - 1].n2w; e.res_beta = h->L[l - 1].n2b; }
            if ((rc = snb_gemm_launch(&P->out[l], EPI_BIAS_BF16, &e, h->num_sms, s))) return rc;
            memset(&e, 0, sizeof(e));                       // linear1 on LN1(z1), folded
            e.bias = h->L[l].b1_f; e.fold = 1; e.colsum = h->L[l].cs_1; e.stats_in = h->stats1;
            if ((rc = snb_gemm_launch(&P->ff1_f[l], EPI_BIAS_RELU_BF16, &e, h->num_sms, s))) return rc;
            memset(&e, 0, sizeof(e));                       // linear2: z2 = ff W_2^T + b_2 + LN1(z1)
            e.bias = h->L[l].b2; e.res = 2; e.resid = h->pre; e.res_stats = h->stats1; e.res_gamma = h->L[l].n1w; e.res_beta = h->L[l].n1b;
            e.stats_out = h->stats2;
            if ((rc = snb_gemm_launch(&P->ff2_f[l], EPI_BIAS_BF16, &e, h->num_sms, s))) return rc;
        }
        memset(&e, 0, sizeof(e));                           // concat3 on LN2(z2) of the last layer, folded
        e.bias = h->bc3_f; e.fold = 1; e.colsum = h->cs_c3; e.stats_in = h->stats2;
        e.gate = h->gate + 512; e.hbias = h->hb + 512; e.tab_ld = HYPER_LD; e.tok_per_env = P->N; e.T = h->T; e.A = P->A;
        if ((rc = snb_gemm_launch(&P->c3_f, EPI_CSL_BF16, &e, h->num_sms, s))) return rc;
    } else {
    for (int l = 0; l < NL; ++l) {
        memset(&e, 0, sizeof(e));
        e.bias = h->L[l].bqkv;
        if ((rc = snb_gemm_launch(&P->qkv[l], EPI_BIAS_BF16, &e, h->num_sms, s))) return rc;
        if (h->joint) rc = attn_v1() ? snb_attn_launch(&P->attn, s) : snb_attn2_launch(&P->attn2, h->att, s);
        else rc = snb_attn_small_launch(h->qkv, h->att, M / h->T, h->T, s);
        if (rc) return rc;
        e.bias = h->L[l].bo;
        if ((rc = snb_gemm_launch(&P->out[l], EPI_BIAS_BF16, &e, h->num_sms, s))) return rc;
        if ((rc = snb_k_layernorm(h->pre, h->h, h->L[l].n1w, h->L[l].n1b, h->y, M, s))) return rc;   // y = LN1(h + attn)
        e.bias = h->L[l].b1;
        if ((rc = snb_gemm_launch(&P->ff1[l], EPI_BIAS_RELU_BF16, &e, h->num_sms, s))) return rc;
        e.bias = h->L[l].b2;
        if ((rc = snb_gemm_launch(&P->ff2[l], EPI_BIAS_BF16, &e, h->num_sms, s))) return rc;
        if ((rc = snb_k_layernorm(h->pre, h->y, h->L[l].n2w, h->L[l].n2b, h->h, M, s))) return rc;   // h = LN2(y + ff)
    }
    memset(&e, 0, sizeof(e));
    e.bias = h->c3_b; e.gate = h->gate + 512; e.hbias = h->hb + 512; e.tab_ld = HYPER_LD;
    e.tok_per_env = P->N; e.T = h->T; e.A = P->A;
    if ((rc = snb_gemm_launch(&P->c3, EPI_CSL_BF16, &e, h->num_sms, s))) return rc;
    }
    memset(&e, 0, sizeof(e));
    e.tab_ld = HYPER_LD; e.tok_per_env = P->N; e.T = h->T; e.A = P->A;
    e.bias = h->c4_b; e.gate = h->gate + 768; e.hbias = h->hb + 768;
    if ((rc = snb_gemm_launch(&P->c4, EPI_CSL_BF16, &e, h->num_sms, s))) return rc;
    // DDIM coefficients in fp32 like torch: (1 - ab).sqrt(), ab.sqrt(), ab_next.sqrt(), (1 - ab_next).sqrt()
    const float ab = h->alpha_bars[t], abn = h->alpha_bars[t_next];
    return snb_k_tail_ddim(h->t4, h->lin_w, h->lin_b, h->gate + 896, h->hb + 896, HYPER_LD, x_in, x_next, eps_out, M, P->N, h->T,
                           P->A, sqrtf(1.0f - ab), sqrtf(ab), sqrtf(abn), sqrtf(1.0f - abn), s);
}

} // namespace

namespace {

int get_x_plans(SnbJmid *h, int n_env, int A, Plans **out)
{
    const std::pair<int, int> key(n_env, A);
    auto it = h->x_plans.find(key);
    if (it != h->x_plans.end()) { *out = &it->second; return SNB_OK; }
    Plans p;
    p.n_env = n_env; p.A = A; p.N = A * h->S * h->T; p.M = n_env * p.N;
    int rc = 0;
    for (int l = 0; l < NL && !rc; ++l) {
        rc = snb_gemm_plan(&p.qkv[l], h->x_a6, h->x_wqkv[l], h->x_qkv, 1, p.M, 3 * D, 6 * D);
        if (!rc) rc = snb_gemm_plan(&p.out[l], h->x_a6, h->x_wo[l], h->x_pre, 1, p.M, D, 6 * D);
        if (!rc) rc = snb_gemm_plan(&p.ff1[l], h->x_a6, h->x_w1[l], h->x_ff, 1, p.M, DFF, 6 * D);
        if (!rc) rc = snb_gemm_plan(&p.ff2[l], h->x_a6, h->x_w2[l], h->x_pre, 1, p.M, D, 6 * DFF);
    }
    if (!rc) rc = snb_gemm_plan(&p.c3, h->x_a6, h->x_wc3, h->x_t3, 1, p.M, 256, 6 * D);
    if (!rc) rc = snb_gemm_plan(&p.c4, h->x_a6, h->x_wc4, h->x_t4, 1, p.M, 128, 6 * 256);
    if (rc) return rc;
    h->x_plans[key] = p;
    *out = &h->x_plans[key];
    return SNB_OK;
}

// fp32-class forward (diffusion.py:173-209): same dataflow as net_forward, fp32 activations, split-bf16 GEMMs
int net_forward_x(SnbJmid *h, Plans *P, const float *x_in, float *x_next, float *eps_out, int t, int t_next, cudaStream_t s)
{
    const int M = P->M, n_ba = P->n_env * P->A;
    int rc = snb_k_hyper_iter(h->hyper, h->gc, h->bc, h->gate, h->hb, n_ba, h->betas[t], s);
    if (rc) return rc;
    if ((rc = snb_x_embed(x_in, h->c1_w, h->c1_b, h->gate, h->hb, h->pe, h->x_h, M, P->N, h->T, P->A, s))) return rc;
    GemmEpi e;
    auto gemm = [&](const GemmPlan *pl, const float *src, int K, int relu, const float *bias) -> int {
        int r = snb_x_split3(src, h->x_a6, (size_t)M, K, 0, relu, s);
        if (r) return r;
        memset(&e, 0, sizeof(e));
        e.bias = bias;
        return snb_gemm_launch(pl, EPI_BIAS_F32, &e, h->num_sms, s);
    };
    for (int l = 0; l < NL; ++l) {
        if ((rc = gemm(&P->qkv[l], h->x_h, D, 0, h->L[l].bqkv))) return rc;
        if ((rc = h->joint ? snb_x_attention(h->x_qkv, h->x_att, P->n_env, P->N, s) : snb_x_attention(h->x_qkv, h->x_att, M / h->T, h->T, s))) return rc;
        if ((rc = gemm(&P->out[l], h->x_att, D, 0, h->L[l].bo))) return rc;
        if ((rc = snb_x_layernorm(h->x_pre, h->x_h, h->L[l].n1w, h->L[l].n1b, h->x_y, M, s))) return rc;
        if ((rc = gemm(&P->ff1[l], h->x_y, D, 0, h->L[l].b1))) return rc;
        if ((rc = gemm(&P->ff2[l], h->x_ff, DFF, 1, h->L[l].b2))) return rc;          // ReLU folded into the split of linear1's output
        if ((rc = snb_x_layernorm(h->x_pre, h->x_y, h->L[l].n2w, h->L[l].n2b, h->x_h, M, s))) return rc;
    }
    if ((rc = gemm(&P->c3, h->x_h, D, 0, h->c3_b))) return rc;
    if ((rc = snb_x_csl_apply(h->x_t3, h->gate + 512, h->hb + 512, (size_t)M, 256, P->N, h->T, P->A, s))) return rc;
    if ((rc = gemm(&P->c4, h->x_t3, 256, 0, h->c4_b))) return rc;
    if ((rc = snb_x_csl_apply(h->x_t4, h->gate + 768, h->hb + 768, (size_t)M, 128, P->N, h->T, P->A, s))) return rc;
    const float ab = h->alpha_bars[t], abn = h->alpha_bars[t_next];
    return snb_x_tail_ddim(h->x_t4, h->lin_w, h->lin_b, h->gate + 896, h->hb + 896, x_in, x_next, eps_out, M, P->N, h->T, P->A,
                           sqrtf(1.0f - ab), sqrtf(ab), sqrtf(abn), sqrtf(1.0f - abn), s);
}

} // namespace

extern "C" int snb_jmid_set_precision(SnbJmid *h, int32_t precision, void *stream)
{
    SNB_REQUIRE(h, SNB_EINVAL, "snb_jmid_set_precision: NULL handle");
    SNB_REQUIRE(precision == SNB_PREC_BF16 || precision == SNB_PREC_FP32X, SNB_EINVAL, "snb_jmid_set_precision: unknown precision %d", precision);
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == SNB_PREC_FP32X && !h->x_a6) {
        // first use: split the weights, allocate fp32 activations for a small chunk (this is a parity instrument: 12 KB of split
        // operand per token row)
        const char *ce = getenv("SNB_JMID_X_CHUNK");
        int chunk = ce ? atoi(ce) : 8;
        if (chunk < 1) chunk = 1;
        h->x_chunk_envs = chunk < h->max_envs ? chunk : h->max_envs;
        const size_t Mc = (((size_t)h->x_chunk_envs * h->N + 127) / 128) * 128;
        int rc = 0;
#define TRY(x) do { if (!rc) rc = (x); } while (0)
        for (int l = 0; l < NL; ++l) {
            TRY(dev_alloc(h, &h->x_wqkv[l], (size_t)3 * D * 6 * D)); TRY(dev_alloc(h, &h->x_wo[l], (size_t)D * 6 * D));
            TRY(dev_alloc(h, &h->x_w1[l], (size_t)DFF * 6 * D)); TRY(dev_alloc(h, &h->x_w2[l], (size_t)D * 6 * DFF));
            TRY(snb_x_split3(h->src_w[l][0], h->x_wqkv[l], 3 * D, D, 1, 0, s)); TRY(snb_x_split3(h->src_w[l][1], h->x_wo[l], D, D, 1, 0, s));
            TRY(snb_x_split3(h->src_w[l][2], h->x_w1[l], DFF, D, 1, 0, s)); TRY(snb_x_split3(h->src_w[l][3], h->x_w2[l], D, DFF, 1, 0, s));
        }
        TRY(dev_alloc(h, &h->x_wc3, (size_t)256 * 6 * D)); TRY(dev_alloc(h, &h->x_wc4, (size_t)128 * 6 * 256));
        TRY(snb_x_split3(h->src_c3, h->x_wc3, 256, D, 1, 0, s)); TRY(snb_x_split3(h->src_c4, h->x_wc4, 128, 256, 1, 0, s));
        TRY(dev_alloc(h, &h->x_a6, Mc * 6 * DFF));
        TRY(dev_alloc(h, &h->x_h, Mc * D)); TRY(dev_alloc(h, &h->x_y, Mc * D)); TRY(dev_alloc(h, &h->x_qkv, Mc * 3 * D));
        TRY(dev_alloc(h, &h->x_att, Mc * D)); TRY(dev_alloc(h, &h->x_ff, Mc * DFF)); TRY(dev_alloc(h, &h->x_pre, Mc * D));
        TRY(dev_alloc(h, &h->x_t3, Mc * 256)); TRY(dev_alloc(h, &h->x_t4, Mc * 128));
#undef TRY
        if (rc) return rc;
    }
    h->precision = precision;
    return SNB_OK;
}

extern "C" double snb_jmid_flops_per_iter(int32_t A, int32_t S, int32_t T, int32_t joint)
{
    const double N = (double)A * S * T, R = (double)A * S;
    const double dense = 12913152.0 * N;
    return dense + (joint ? 6144.0 * N * N : 6144.0 * R * T * T);
}

extern "C" int snb_jmid_create(SnbJmid **out, const SnbJmidWeights *w, int32_t max_envs, int32_t A, int32_t S, int32_t T,
                               int32_t joint, void *stream)
{
    SNB_REQUIRE(out && w, SNB_EINVAL, "snb_jmid_create: NULL argument");
    SNB_REQUIRE(max_envs >= 1 && A >= 1 && S >= 1 && T >= 1 && T <= 24, SNB_EINVAL, "snb_jmid_create: bad sizes (T <= 24 = pos_emb max_len)");
    int dev = 0, cc_major = 0;
    SNB_CUDA_TRY(cudaGetDevice(&dev));
    SNB_CUDA_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    SNB_REQUIRE(cc_major == 10, SNB_ECUDA, "snb_jmid_create: the denoiser kernels are sm_100a only (device is sm_%d)", cc_major * 10);
    cudaStream_t s = (cudaStream_t)stream;
    SnbJmid *h = new SnbJmid();
    h->A = A; h->S = S; h->T = T; h->joint = joint ? 1 : 0; h->max_envs = max_envs; h->N = A * S * T;
    SNB_CUDA_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev));
    const char *ce = getenv("SNB_JMID_CHUNK");
    int chunk = ce ? atoi(ce) : 512;   // envs per chunk: 512 x 1600 tokens = 8 GB of activations; +2.5 % over 128 on the same box (r01)
    if (chunk < 1) chunk = 1;
    h->chunk_envs = chunk < max_envs ? chunk : max_envs;
    int rc = 0;
#define TRY(x) do { if (!rc) rc = (x); } while (0)
    const SnbCslWeights *csl[4] = {&w->concat1, &w->concat3, &w->concat4, &w->linear};
    const int douts[4] = {512, 256, 128, 2};
    for (int i = 0; i < 4; ++i) {
        float *gw = nullptr, *gb = nullptr, *bw = nullptr;
        TRY(dup_f32(h, &gw, (const float *)csl[i]->hyper_gate_w, (size_t)douts[i] * 259, s));
        TRY(dup_f32(h, &gb, (const float *)csl[i]->hyper_gate_b, (size_t)douts[i], s));
        TRY(dup_f32(h, &bw, (const float *)csl[i]->hyper_bias_w, (size_t)douts[i] * 259, s));
        h->hyper[i].gate_w = gw; h->hyper[i].gate_b = gb; h->hyper[i].bias_w = bw; h->hyper[i].dout = douts[i];
    }
    TRY(dup_f32(h, &h->c1_w, (const float *)w->concat1.layer_w, 512 * 2, s));
    TRY(dup_f32(h, &h->c1_b, (const float *)w->concat1.layer_b, 512, s));
    {
        float *c3 = nullptr, *c4 = nullptr;
        TRY(dup_f32(h, &c3, (const float *)w->concat3.layer_w, 256 * 512, s)); TRY(dup_f32(h, &c4, (const float *)w->concat4.layer_w, 128 * 256, s));
        h->src_c3 = c3; h->src_c4 = c4;
    }
    TRY(dup_bf16(h, &h->wc3, (const float *)w->concat3.layer_w, 256 * 512, s));
    TRY(dup_f32(h, &h->c3_b, (const float *)w->concat3.layer_b, 256, s));
    TRY(dup_bf16(h, &h->wc4, (const float *)w->concat4.layer_w, 128 * 256, s));
    TRY(dup_f32(h, &h->c4_b, (const float *)w->concat4.layer_b, 128, s));
    TRY(dup_f32(h, &h->lin_w, (const float *)w->linear.layer_w, 2 * 128, s));
    TRY(dup_f32(h, &h->lin_b, (const float *)w->linear.layer_b, 2, s));
    TRY(dup_f32(h, &h->pe, (const float *)w->pos_emb, (size_t)T * 512, s));
    for (int l = 0; l < NL; ++l) {
        const SnbEncLayerWeights &lw = w->layers[l];
        {   // fp32 copies of the four matrices: the fp32x path splits them on first use (27 MB, only touched then)
            float *c0 = nullptr, *c1 = nullptr, *c2 = nullptr, *c3 = nullptr;
            TRY(dup_f32(h, &c0, (const float *)lw.in_proj_w, (size_t)3 * D * D, s)); TRY(dup_f32(h, &c1, (const float *)lw.out_proj_w, (size_t)D * D, s));
            TRY(dup_f32(h, &c2, (const float *)lw.lin1_w, (size_t)DFF * D, s)); TRY(dup_f32(h, &c3, (const float *)lw.lin2_w, (size_t)D * DFF, s));
            h->src_w[l][0] = c0; h->src_w[l][1] = c1; h->src_w[l][2] = c2; h->src_w[l][3] = c3;
        }
        TRY(dup_bf16(h, &h->L[l].wqkv, (const float *)lw.in_proj_w, (size_t)3 * D * D, s));
        TRY(dup_f32(h, &h->L[l].bqkv, (const float *)lw.in_proj_b, 3 * D, s));
        TRY(dup_bf16(h, &h->L[l].wo, (const float *)lw.out_proj_w, (size_t)D * D, s));
        TRY(dup_f32(h, &h->L[l].bo, (const float *)lw.out_proj_b, D, s));
        TRY(dup_bf16(h, &h->L[l].w1, (const float *)lw.lin1_w, (size_t)DFF * D, s));
        TRY(dup_f32(h, &h->L[l].b1, (const float *)lw.lin1_b, DFF, s));
        TRY(dup_bf16(h, &h->L[l].w2, (const float *)lw.lin2_w, (size_t)D * DFF, s));
        TRY(dup_f32(h, &h->L[l].b2, (const float *)lw.lin2_b, D, s));
        TRY(dup_f32(h, &h->L[l].n1w, (const float *)lw.norm1_w, D, s));
        TRY(dup_f32(h, &h->L[l].n1b, (const float *)lw.norm1_b, D, s));
        TRY(dup_f32(h, &h->L[l].n2w, (const float *)lw.norm2_w, D, s));
        TRY(dup_f32(h, &h->L[l].n2b, (const float *)lw.norm2_b, D, s));
    }
    {
        const char *fe = getenv("SNB_LN_FOLD");
        h->ln_fold = fe ? atoi(fe) : 1;
    }
    if (h->ln_fold) {
        for (int l = 0; l < NL; ++l) {
            const SnbEncLayerWeights &lw = w->layers[l];
            TRY(dev_alloc(h, &h->L[l].w1_f, (size_t)DFF * D)); TRY(dev_alloc(h, &h->L[l].cs_1, DFF)); TRY(dev_alloc(h, &h->L[l].b1_f, DFF));
            TRY(snb_k_fold_ln((const float *)lw.lin1_w, (const float *)lw.norm1_w, (const float *)lw.norm1_b, (const float *)lw.lin1_b,
                              h->L[l].w1_f, h->L[l].cs_1, h->L[l].b1_f, DFF, D, s));
            if (l > 0) {
                const SnbEncLayerWeights &pw = w->layers[l - 1];
                TRY(dev_alloc(h, &h->L[l].wqkv_f, (size_t)3 * D * D)); TRY(dev_alloc(h, &h->L[l].cs_qkv, 3 * D)); TRY(dev_alloc(h, &h->L[l].bqkv_f, 3 * D));
                TRY(snb_k_fold_ln((const float *)lw.in_proj_w, (const float *)pw.norm2_w, (const float *)pw.norm2_b, (const float *)lw.in_proj_b,
                                  h->L[l].wqkv_f, h->L[l].cs_qkv, h->L[l].bqkv_f, 3 * D, D, s));
            }
        }
        const SnbEncLayerWeights &last = w->layers[NL - 1];
        TRY(dev_alloc(h, &h->wc3_f, (size_t)256 * D)); TRY(dev_alloc(h, &h->cs_c3, 256)); TRY(dev_alloc(h, &h->bc3_f, 256));
        TRY(snb_k_fold_ln((const float *)w->concat3.layer_w, (const float *)last.norm2_w, (const float *)last.norm2_b,
                          (const float *)w->concat3.layer_b, h->wc3_f, h->cs_c3, h->bc3_f, 256, D, s));
    }
    const size_t Mc = (size_t)h->chunk_envs * h->N;
    const size_t Mp = ((Mc + 127) / 128) * 128; // padded so that tensor-map boxes never leave the allocation
    if (h->ln_fold) { TRY(dev_alloc(h, &h->stats1, Mp * 4)); TRY(dev_alloc(h, &h->stats2, Mp * 4)); }
    TRY(dev_alloc(h, &h->h, Mp * D));
    TRY(dev_alloc(h, &h->y, Mp * D));
    TRY(dev_alloc(h, &h->qkv, Mp * 3 * D));
    TRY(dev_alloc(h, &h->att, Mp * D));
    TRY(dev_alloc(h, &h->ff, Mp * DFF));
    TRY(dev_alloc(h, &h->t3, Mp * 256));
    TRY(dev_alloc(h, &h->t4, Mp * 128));
    TRY(dev_alloc(h, &h->pre, Mp * D));
    TRY(dev_alloc(h, &h->xa, Mp * 2));
    TRY(dev_alloc(h, &h->xb, Mp * 2));
    const size_t nba = (size_t)h->chunk_envs * A;
    TRY(dev_alloc(h, &h->gc, nba * HYPER_LD));
    TRY(dev_alloc(h, &h->bc, nba * HYPER_LD));
    TRY(dev_alloc(h, &h->gate, nba * HYPER_LD));
    TRY(dev_alloc(h, &h->hb, nba * HYPER_LD));
    TRY(dev_alloc(h, &h->ctx_stage, nba * 256));
#undef TRY
    {
        const char *ge = getenv("SNB_JMID_GRAPH");
        h->use_graphs = ge ? atoi(ge) : 1;
    }
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(h->betas, w->betas, sizeof(float) * 101, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->alpha_bars, w->alpha_bars, sizeof(float) * 101, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { snb_set_error("snb_jmid_create: %s", cudaGetErrorString(e)); rc = SNB_ECUDA; }
    }
    if (rc) { snb_jmid_destroy(h); return rc; }
    *out = h;
    return SNB_OK;
}

extern "C" int snb_jmid_destroy(SnbJmid *h)
{
    if (!h) return SNB_OK;
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (void *p : h->allocs) cudaFree(p);
    cudaFree(h->d_ctx); cudaFree(h->d_xT); cudaFree(h->d_p0); cudaFree(h->d_vel); cudaFree(h->d_pos);
    delete h;
    return SNB_OK;
}

namespace {
// hyper_ctx + the whole DDIM loop for the chunk staged in h->ctx_stage / h->xa; the result ends in *result
int chunk_sequence(SnbJmid *h, Plans *P, int n_steps, cudaStream_t s, float **result)
{
    const int stride = 100 / n_steps; // int(100 / step), diffusion.py:507
    int rc = snb_k_hyper_ctx(h->hyper, h->ctx_stage, h->gc, h->bc, P->n_env * P->A, s);
    if (rc) return rc;
    float *cur = h->xa, *nxt = h->xb;
    for (int t = 100; t > 0; t -= stride) {
        if ((rc = (h->precision == SNB_PREC_FP32X ? net_forward_x : net_forward)(h, P, cur, nxt, nullptr, t, t - stride, s))) return rc;
        float *tmp = cur; cur = nxt; nxt = tmp;
    }
    *result = cur;
    return SNB_OK;
}
} // namespace

extern "C" int snb_jmid_denoise_agents(SnbJmid *h, const float *ctx, const float *x_T, float *out_vel, int32_t B, int32_t A,
                                       int32_t n_steps, void *stream)
{
    SNB_REQUIRE(h && ctx && x_T && out_vel, SNB_EINVAL, "snb_jmid_denoise: NULL argument");
    SNB_REQUIRE(B >= 0 && n_steps >= 1 && n_steps <= 100, SNB_EINVAL, "snb_jmid_denoise: bad B / n_steps");
    SNB_REQUIRE(A >= 1 && A <= h->A, SNB_EINVAL, "snb_jmid_denoise: A=%d outside [1, %d] (the handle's agent capacity)", A, h->A);
    cudaStream_t s = (cudaStream_t)stream;
    const int stride = 100 / n_steps;
    // t = 100, 100 - stride, ... must land on 0: the reference returns traj[0] and raises KeyError otherwise (diffusion.py:507-537)
    SNB_REQUIRE(100 % stride == 0, SNB_EINVAL,
                "snb_jmid_denoise: step_size=%d gives stride int(100/%d)=%d, which does not divide the 100 diffusion steps "
                "(the reference raises KeyError: 0 on traj[0])", n_steps, n_steps, stride);
    const int n_iter = (100 + stride - 1) / stride;
    const int N = A * h->S * h->T;
    const bool fx = h->precision == SNB_PREC_FP32X;
    // a chunk is bounded by ROWS (chunk_envs * tokens at full A): fewer agents per env -> more envs per chunk
    int chunk = (int)(((size_t)(fx ? h->x_chunk_envs : h->chunk_envs) * h->N) / N);
    if (chunk < 1) chunk = 1;
    for (int e0 = 0; e0 < B; e0 += chunk) {
        const int ne = (B - e0) < chunk ? (B - e0) : chunk;
        Plans *P = nullptr;
        int rc = fx ? get_x_plans(h, ne, A, &P) : get_plans(h, ne, A, &P);
        if (rc) return rc;
        const size_t M = (size_t)P->M;
        SNB_CUDA_TRY(cudaMemcpyAsync(h->ctx_stage, ctx + (size_t)e0 * A * 256, (size_t)ne * A * 256 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        SNB_CUDA_TRY(cudaMemcpyAsync(h->xa, x_T + (size_t)e0 * N * 2, M * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        float *result = (n_iter & 1) ? h->xb : h->xa;
        const int64_t key = ((int64_t)ne << 24) | ((int64_t)A << 8) | (int64_t)n_steps;
        auto git = fx ? h->graphs.end() : h->graphs.find(key);
        if (git != h->graphs.end()) {
            SNB_CUDA_TRY(cudaGraphLaunch(git->second, s));
            snb_count_launch(1 + n_iter * (h->ln_fold ? 20 : 26));   // kernels inside the replayed graph
        } else if (!fx && h->use_graphs && s != nullptr && ++h->graph_uses[key] >= 2) {
            // second use of this shape: capture the sequence once, then replay it for every later chunk
            cudaGraph_t graph = nullptr;
            SNB_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            float *res2 = nullptr;
            rc = chunk_sequence(h, P, n_steps, s, &res2);
            snb_count_launch(-(1 + n_iter * (h->ln_fold ? 20 : 26))); // launches recorded during capture did not execute
            cudaError_t ce = cudaStreamEndCapture(s, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) { snb_set_error("snb_jmid_denoise: graph capture failed: %s", cudaGetErrorString(ce)); return SNB_ECUDA; }
            cudaGraphExec_t exec = nullptr;
            ce = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { snb_set_error("snb_jmid_denoise: graph instantiate failed: %s", cudaGetErrorString(ce)); return SNB_ECUDA; }
            h->graphs[key] = exec;
            SNB_CUDA_TRY(cudaGraphLaunch(exec, s));
            snb_count_launch(1 + n_iter * (h->ln_fold ? 20 : 26));
        } else {
            float *res2 = nullptr;
            if ((rc = chunk_sequence(h, P, n_steps, s, &res2))) return rc;
            result = res2;
        }
        SNB_CUDA_TRY(cudaMemcpyAsync(out_vel + (size_t)e0 * N * 2, result, M * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return SNB_OK;
}

extern "C" int snb_jmid_denoise(SnbJmid *h, const float *ctx, const float *x_T, float *out_vel, int32_t B, int32_t n_steps, void *stream)
{
    SNB_REQUIRE(h, SNB_EINVAL, "snb_jmid_denoise: NULL handle");
    return snb_jmid_denoise_agents(h, ctx, x_T, out_vel, B, h->A, n_steps, stream);
}

extern "C" int snb_jmid_dims(const SnbJmid *h, int32_t *A, int32_t *S, int32_t *T, int32_t *joint)
{
    SNB_REQUIRE(h, SNB_EINVAL, "snb_jmid_dims: NULL handle");
    if (A) *A = h->A;
    if (S) *S = h->S;
    if (T) *T = h->T;
    if (joint) *joint = h->joint;
    return SNB_OK;
}

extern "C" int snb_jmid_eps(SnbJmid *h, const float *ctx, const float *x_t, float *eps, int32_t B, int32_t t, void *stream)
{
    SNB_REQUIRE(h && ctx && x_t && eps, SNB_EINVAL, "snb_jmid_eps: NULL argument");
    SNB_REQUIRE(t >= 1 && t <= 100, SNB_EINVAL, "snb_jmid_eps: t out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const bool fx = h->precision == SNB_PREC_FP32X;
    const int ce = fx ? h->x_chunk_envs : h->chunk_envs;
    for (int e0 = 0; e0 < B; e0 += ce) {
        const int ne = (B - e0) < ce ? (B - e0) : ce;
        Plans *P = nullptr;
        int rc = fx ? get_x_plans(h, ne, h->A, &P) : get_plans(h, ne, h->A, &P);
        if (rc) return rc;
        if ((rc = snb_k_hyper_ctx(h->hyper, ctx + (size_t)e0 * h->A * 256, h->gc, h->bc, ne * h->A, s))) return rc;
        if ((rc = (fx ? net_forward_x : net_forward)(h, P, x_t + (size_t)e0 * h->N * 2, nullptr, eps + (size_t)e0 * h->N * 2, t, t - 1, s))) return rc;
    }
    return SNB_OK;
}

extern "C" int snb_jmid_integrate(const float *vel, const float *p0, float *pos, int32_t B, int32_t S, int32_t A, int32_t T,
                                  float dt, void *stream)
{
    SNB_REQUIRE(vel && p0 && pos, SNB_EINVAL, "snb_jmid_integrate: NULL argument");
    if (B == 0) return SNB_OK;
    return snb_k_integrate(vel, p0, pos, B, S, A, T, dt, (cudaStream_t)stream);
}

extern "C" int snb_jmid_predict_host(SnbJmid *h, const float *ctx_host, const float *x_T_host, const float *p0_host,
                                     float *pos_host, int32_t B, int32_t n_steps, float dt)
{
    SNB_REQUIRE(h && ctx_host && x_T_host && p0_host && pos_host, SNB_EINVAL, "snb_jmid_predict_host: NULL argument");
    SNB_REQUIRE(B >= 1 && B <= h->max_envs, SNB_EINVAL, "snb_jmid_predict_host: B=%d beyond max_envs=%d", B, h->max_envs);
    if (!h->d_ctx) {
        const size_t me = (size_t)h->max_envs;
        SNB_CUDA_TRY(cudaMalloc(&h->d_ctx, me * h->A * 256 * sizeof(float)));
        SNB_CUDA_TRY(cudaMalloc(&h->d_xT, me * h->N * 2 * sizeof(float)));
        SNB_CUDA_TRY(cudaMalloc(&h->d_p0, me * h->A * 2 * sizeof(float)));
        SNB_CUDA_TRY(cudaMalloc(&h->d_vel, me * h->N * 2 * sizeof(float)));
        SNB_CUDA_TRY(cudaMalloc(&h->d_pos, me * h->N * 2 * sizeof(float)));
    }
    if (!h->own_stream) SNB_CUDA_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    cudaStream_t s = h->own_stream;
    SNB_CUDA_TRY(cudaMemcpyAsync(h->d_ctx, ctx_host, (size_t)B * h->A * 256 * sizeof(float), cudaMemcpyHostToDevice, s));
    SNB_CUDA_TRY(cudaMemcpyAsync(h->d_xT, x_T_host, (size_t)B * h->N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    SNB_CUDA_TRY(cudaMemcpyAsync(h->d_p0, p0_host, (size_t)B * h->A * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = snb_jmid_denoise(h, h->d_ctx, h->d_xT, h->d_vel, B, n_steps, s);
    if (rc) return rc;
    rc = snb_jmid_integrate(h->d_vel, h->d_p0, h->d_pos, B, h->S, h->A, h->T, dt, s);
    if (rc) return rc;
    SNB_CUDA_TRY(cudaMemcpyAsync(pos_host, h->d_pos, (size_t)B * h->N * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    SNB_CUDA_TRY(cudaStreamSynchronize(s));
    return SNB_OK;
}

extern "C" int snb_jmid_gemm_bf16(const void *A, const void *W, const float *bias, void *out, int32_t M, int32_t N, int32_t K,
                                  int32_t epi, void *stream)
{
    SNB_REQUIRE(A && W && bias && out, SNB_EINVAL, "snb_jmid_gemm_bf16: NULL argument");
    SNB_REQUIRE(epi >= 0 && epi <= 2, SNB_EINVAL, "snb_jmid_gemm_bf16: bad epilogue");
    int dev = 0, sms = 0;
    SNB_CUDA_TRY(cudaGetDevice(&dev));
    SNB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GemmPlan p;
    int rc = snb_gemm_plan(&p, (const bf16 *)A, (const bf16 *)W, out, epi == 2, M, N, K);
    if (rc) return rc;
    GemmEpi e;
    memset(&e, 0, sizeof(e));
    e.bias = bias;
    return snb_gemm_launch(&p, epi, &e, sms, (cudaStream_t)stream);
}

extern "C" int snb_jmid_attention(const void *qkv, void *out, int32_t n_env, int32_t n_tok, void *stream)
{
    SNB_REQUIRE(qkv && out, SNB_EINVAL, "snb_jmid_attention: NULL argument");
    if (attn_v1()) {
        AttnPlan p;
        int rc = snb_attn_plan(&p, (const bf16 *)qkv, (bf16 *)out, n_env, n_tok);
        if (rc) return rc;
        return snb_attn_launch(&p, (cudaStream_t)stream);
    }
    Attn2Plan p;
    int rc = snb_attn2_plan(&p, (const bf16 *)qkv, n_env, n_tok);
    if (rc) return rc;
    return snb_attn2_launch(&p, (bf16 *)out, (cudaStream_t)stream);
}
