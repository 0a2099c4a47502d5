// csrc/jmid_fp32x.cu -- the fp32-class path of the JMID noise network (precision "fp32x"), used to separate rounding error from
// algorithmic error in the parity suite (SURVEY 8d gate: eps <= 1e-4 vs models/diffusion.py:173-209, which computes in fp32).
//
// Activations stay fp32 in HBM.  Every nn.Linear runs on the SAME tcgen05 GEMM kernel as the bf16 path: an fp32 value is split into
// three bf16 pieces a = hi + mid + lo (24 mantissa bits), and the six significant partial products
//     hi*hi + hi*mid + mid*hi + hi*lo + lo*hi + mid*mid          (dropped terms <= 2^-24 |a||w|)
// are laid out along K (K' = 6K), accumulated in fp32 in TMEM:  A' = [hi hi mid hi lo mid],  W' = [hi mid hi lo hi mid].
// Attention, LayerNorm, ConcatSquash and the tail run in fp32 on the CUDA cores (FFMA, expf).  6x the tensor FLOPs and a SIMT
// attention: a parity instrument, not the product path (bench.py reports its throughput beside the bf16 line).
#include "jmid_internal.h"

namespace {

__device__ __forceinline__ void split3(float a, bf16 &hi, bf16 &mid, bf16 &lo)
{
    hi = __float2bfloat16_rn(a);
    const float r1 = a - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

// src [rows, K] fp32 -> dst [rows, 6K] bf16; mode 0: activation order [hi hi mid hi lo mid], mode 1: weight order [hi mid hi lo hi mid]
__global__ void split3_kernel(const float *__restrict__ src, bf16 *__restrict__ dst, size_t rows, int K, int mode, int relu)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * (size_t)K) return;
    const size_t r = i / K;
    const int k = (int)(i - r * K);
    float a = src[i];
    if (relu) a = fmaxf(a, 0.0f);
    bf16 hi, mid, lo;
    split3(a, hi, mid, lo);
    bf16 *d = dst + r * (size_t)(6 * K) + k;
    if (mode == 0) { d[0] = hi; d[K] = hi; d[2 * K] = mid; d[3 * K] = hi; d[4 * K] = lo; d[5 * K] = mid; }
    else { d[0] = hi; d[K] = mid; d[2 * K] = hi; d[3 * K] = lo; d[4 * K] = hi; d[5 * K] = mid; }
}

// concat1 (2 -> 512) * gate + hyper-bias + positional encoding (common.py:65-72, diffusion.py:183-185), fp32 out
__global__ void embed_f32_kernel(const float *__restrict__ x, const float *__restrict__ w1, const float *__restrict__ b1,
                                 const float *__restrict__ gate, const float *__restrict__ hb, const float *__restrict__ pe,
                                 float *__restrict__ h, int n_tok_total, int tok_per_env, int T, int A)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_tok_total * 512) return;
    const int col = (int)(i & 511);
    const size_t tok = i >> 9;
    const int b = (int)(tok / tok_per_env);
    const int rem = (int)(tok - (size_t)b * tok_per_env);
    const int r = rem / T, tau = rem - r * T;
    const size_t ba = (size_t)b * A + (r % A);
    const float lin = fmaf(x[2 * tok + 1], w1[2 * col + 1], fmaf(x[2 * tok], w1[2 * col], b1[col]));
    h[i] = fmaf(lin, gate[ba * HYPER_LD + col], hb[ba * HYPER_LD + col]) + pe[(size_t)tau * 512 + col];
}

// out = LayerNorm(512)(pre + resid), all fp32, one warp per row (two-pass statistics like torch)
__global__ void layernorm_f32_kernel(const float *__restrict__ pre, const float *__restrict__ resid, const float *__restrict__ g,
                                     const float *__restrict__ b, float *__restrict__ out, int rows)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float v[16];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const size_t k = (size_t)row * 512 + lane + 32 * i;
        v[i] = pre[k] + resid[k];
        sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / 512.0f);
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq * (1.0f / 512.0f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = lane + 32 * i;
        out[(size_t)row * 512 + c] = (v[i] - mean) * rstd * g[c] + b[c];
    }
}

// in-place ConcatSquash epilogue on a GEMM output that already holds acc + bias: v = v * gate[ba, n] + hbias[ba, n]
__global__ void csl_apply_f32_kernel(float *__restrict__ v, const float *__restrict__ gate, const float *__restrict__ hb, size_t rows, int N,
                                     int tok_per_env, int T, int A)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * (size_t)N) return;
    const size_t m = i / N;
    const int n = (int)(i - m * N);
    const int b = (int)(m / tok_per_env);
    const int r = (int)(m - (size_t)b * tok_per_env) / T;
    const size_t ba = (size_t)b * A + (r % A);
    v[i] = fmaf(v[i], gate[ba * HYPER_LD + n], hb[ba * HYPER_LD + n]);
}

// multi-head self-attention in fp32: one unmasked sequence of n_tok tokens per environment (JMID, diffusion.py:196-204), or with
// seq_len = T independent sequences (iMID).  Block = 32 queries x 4 threads (32 head dims each); K / V tiles of 64 keys in shared memory.
constexpr int FA_QT = 32, FA_KT = 64;
__global__ void __launch_bounds__(128) attn_f32_kernel(const float *__restrict__ qkv, float *__restrict__ out, int seq_len, float scale)
{
    extern __shared__ float sm[];
    float *sK = sm, *sV = sm + FA_KT * 128;
    const int seq = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * FA_QT;
    const int qi = threadIdx.x >> 2, part = threadIdx.x & 3;          // query within the tile, 32-dim slice of the head
    const int q = q0 + qi;
    const bool live = q < seq_len;
    const float *base = qkv + (size_t)seq * seq_len * 1536 + head * 128;
    float qv[32], o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) { qv[d] = live ? base[(size_t)q * 1536 + part * 32 + d] * scale : 0.0f; o[d] = 0.0f; }
    float m = -INFINITY, l = 0.0f;
    for (int k0 = 0; k0 < seq_len; k0 += FA_KT) {
        const int nk = min(FA_KT, seq_len - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < FA_KT * 32; i += 128) {          // float4 granules: 64 keys x 32 granules
            const int kk = i >> 5, g4 = i & 31;
            float4 kx = make_float4(0.f, 0.f, 0.f, 0.f), vx = kx;
            if (kk < nk) {
                kx = *reinterpret_cast<const float4 *>(base + 512 + (size_t)(k0 + kk) * 1536 + g4 * 4);
                vx = *reinterpret_cast<const float4 *>(base + 1024 + (size_t)(k0 + kk) * 1536 + g4 * 4);
            }
            *reinterpret_cast<float4 *>(sK + kk * 128 + g4 * 4) = kx;
            *reinterpret_cast<float4 *>(sV + kk * 128 + g4 * 4) = vx;
        }
        __syncthreads();
        float sc[FA_KT];
        float tmax = -INFINITY;
#pragma unroll 4
        for (int kk = 0; kk < FA_KT; ++kk) {
            const float *kr = sK + kk * 128 + part * 32;
            float d = 0.0f;
#pragma unroll
            for (int e = 0; e < 32; ++e) d = fmaf(qv[e], kr[e], d);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            sc[kk] = kk < nk ? d : -INFINITY;
            tmax = fmaxf(tmax, sc[kk]);
        }
        const float m_new = fmaxf(m, tmax);
        const float f = expf(m - m_new);                                // 0 on the first tile (m = -inf)
        l *= f;
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] *= f;
#pragma unroll 4
        for (int kk = 0; kk < FA_KT; ++kk) {
            const float p = expf(sc[kk] - m_new);
            l += p;
            const float *vr = sV + kk * 128 + part * 32;
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = fmaf(p, vr[e], o[e]);
        }
        m = m_new;
    }
    if (live) {
        const float inv = 1.0f / l;
        float *dst = out + ((size_t)seq * seq_len + q) * 512 + head * 128 + part * 32;
#pragma unroll
        for (int e = 0; e < 32; ++e) dst[e] = o[e] * inv;
    }
}

// `linear` ConcatSquash 128 -> 2 on the fp32 concat4 output + the DDIM update (diffusion.py:207-209, 524-528); one thread per token
__global__ void tail_ddim_f32_kernel(const float *__restrict__ t4, const float *__restrict__ wl, const float *__restrict__ bl,
                                     const float *__restrict__ gate, const float *__restrict__ hb, const float *__restrict__ x_t,
                                     float *__restrict__ x_next, float *__restrict__ eps_out, int n_tok_total, int tok_per_env, int T, int A,
                                     float c1, float c2, float c3, float c4)
{
    const int tok = blockIdx.x * blockDim.x + threadIdx.x;
    if (tok >= n_tok_total) return;
    float a0 = 0.0f, a1 = 0.0f;
    const float *p = t4 + (size_t)tok * 128;
#pragma unroll 8
    for (int k = 0; k < 128; ++k) { a0 = fmaf(p[k], wl[k], a0); a1 = fmaf(p[k], wl[128 + k], a1); }
    const int b = tok / tok_per_env;
    const int r = (tok - b * tok_per_env) / T;
    const size_t ba = (size_t)b * A + (r % A);
    const float e0 = (a0 + bl[0]) * gate[ba * HYPER_LD] + hb[ba * HYPER_LD];
    const float e1 = (a1 + bl[1]) * gate[ba * HYPER_LD + 1] + hb[ba * HYPER_LD + 1];
    if (eps_out) { eps_out[2 * (size_t)tok] = e0; eps_out[2 * (size_t)tok + 1] = e1; }
    if (x_next) {
        const float x0 = x_t[2 * (size_t)tok], x1 = x_t[2 * (size_t)tok + 1];
        const float p0 = (x0 - e0 * c1) / c2, p1 = (x1 - e1 * c1) / c2;
        x_next[2 * (size_t)tok] = c3 * p0 + c4 * e0;
        x_next[2 * (size_t)tok + 1] = c3 * p1 + c4 * e1;
    }
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

} // namespace

int snb_x_split3(const float *src, bf16 *dst, size_t rows, int K, int weight_order, int relu, cudaStream_t s)
{
    if (rows == 0) return SNB_OK;
    split3_kernel<<<blocks_for(rows * (size_t)K, 256), 256, 0, s>>>(src, dst, rows, K, weight_order, relu);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_x_embed(const float *x, const float *w1, const float *b1, const float *gate, const float *hb, const float *pe, float *h,
                int n_tok_total, int tok_per_env, int T, int A, cudaStream_t s)
{
    embed_f32_kernel<<<blocks_for((size_t)n_tok_total * 512, 256), 256, 0, s>>>(x, w1, b1, gate, hb, pe, h, n_tok_total, tok_per_env, T, A);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_x_layernorm(const float *pre, const float *resid, const float *g, const float *b, float *out, int rows, cudaStream_t s)
{
    layernorm_f32_kernel<<<(rows + 7) / 8, 256, 0, s>>>(pre, resid, g, b, out, rows);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_x_csl_apply(float *v, const float *gate, const float *hb, size_t rows, int N, int tok_per_env, int T, int A, cudaStream_t s)
{
    csl_apply_f32_kernel<<<blocks_for(rows * (size_t)N, 256), 256, 0, s>>>(v, gate, hb, rows, N, tok_per_env, T, A);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_x_attention(const float *qkv, float *out, int n_seq, int seq_len, cudaStream_t s)
{
    static bool attr_set = false;
    const int smem = 2 * FA_KT * 128 * (int)sizeof(float);
    if (!attr_set) {
        SNB_CUDA_TRY(cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    // grid.z is limited to 65535 sequences per launch (iMID has n_env * A * S of them)
    for (int z0 = 0; z0 < n_seq; z0 += 65535) {
        const int nz = n_seq - z0 < 65535 ? n_seq - z0 : 65535;
        dim3 grid((seq_len + FA_QT - 1) / FA_QT, 4, nz);
        attn_f32_kernel<<<grid, 128, smem, s>>>(qkv + (size_t)z0 * seq_len * 1536, out + (size_t)z0 * seq_len * 512, seq_len,
                                                1.0f / sqrtf(128.0f));
        snb_count_launch();
    }
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_x_tail_ddim(const float *t4, const float *wl, const float *bl, const float *gate, const float *hb, const float *x_t, float *x_next,
                    float *eps_out, int n_tok_total, int tok_per_env, int T, int A, float c1, float c2, float c3, float c4, cudaStream_t s)
{
    tail_ddim_f32_kernel<<<blocks_for((size_t)n_tok_total, 128), 128, 0, s>>>(t4, wl, bl, gate, hb, x_t, x_next, eps_out, n_tok_total,
                                                                             tok_per_env, T, A, c1, c2, c3, c4);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
