// csrc/pred_post.cu -- back end of the JMID predictor around the denoiser: noise, bucket gather / scatter, assembly of
// predict_ret_best's outputs (mid_sim_wrapper.py:482-509) and the KDE top-k (get_most_likely_samples, :14-169).
#include <math.h>

#include "pred_internal.h"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller.  The reference draws torch.randn on its device (diffusion.py:499, quirk q2): the stream
// of a different generator cannot be reproduced, parity tests inject the noise instead.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__global__ void noise_kernel(float *__restrict__ out, size_t n, uint64_t seed, uint64_t offset)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // one counter -> 4 normals
    if (q * 4 >= n) return;
    const uint64_t ctr = q + offset;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5a5a5a5au, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    float z[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const float u1 = ((float)c[2 * p] + 1.0f) * 2.3283064365386963e-10f; // (0, 1]
        const float u2 = (float)c[2 * p + 1] * 2.3283064365386963e-10f;
        const float r = sqrtf(-2.0f * logf(fminf(u1, 1.0f)));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z[2 * p] = r * cs; z[2 * p + 1] = r * sn;
    }
    for (int i = 0; i < 4; ++i)
        if (q * 4 + i < n) out[q * 4 + i] = z[i];
}

__global__ void gather_kernel(const int32_t *__restrict__ order, int cnt, int A, int H, int S, int T, const float *__restrict__ ctx,
                              const float *__restrict__ noise, const float *__restrict__ p0, float *__restrict__ ctx_b,
                              float *__restrict__ xT_b, float *__restrict__ p0_b)
{
    const int n_ctx = A * 256, n_x = S * A * T * 2, n_p = A * 2, per = n_ctx + n_x + n_p;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)cnt * per) return;
    const int e = (int)(i / per), r = (int)(i % per);
    const int env = order[e];
    if (r < n_ctx) {
        ctx_b[(size_t)e * n_ctx + r] = ctx[(size_t)env * H * 256 + r]; // slots 0..A-1 are contiguous
    } else if (r < n_ctx + n_x) {
        const int k = r - n_ctx, c = k % (T * 2), a = (k / (T * 2)) % A, s = k / (T * 2 * A);
        xT_b[(size_t)e * n_x + k] = noise[(((size_t)env * S + s) * H + a) * (T * 2) + c];
    } else {
        const int k = r - n_ctx - n_x;
        p0_b[(size_t)e * n_p + k] = p0[(size_t)env * H * 2 + k];
    }
}

// forecasts [B,H,k,T+1,2] fp64; sample slot j of in-cluster agent a  <-  pos_b[e, sel ? sel[e,j] : j, a, :, :]
__global__ void scatter_kernel(const int32_t *__restrict__ order, int cnt, int A, int H, int S, int T, int k,
                               const float *__restrict__ pos_b, const int32_t *__restrict__ sel, const int32_t *__restrict__ ped_ids,
                               double *__restrict__ forecasts)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = A * k * T * 2;
    if (i >= (size_t)cnt * per) return;
    const int e = (int)(i / per), r = (int)(i % per);
    const int c = r % (T * 2), j = (r / (T * 2)) % k, a = r / (T * 2 * k);
    const int env = order[e];
    const int s = sel ? sel[(size_t)e * k + j] : j;
    const int ped = ped_ids[(size_t)env * H + a];
    const float v = pos_b[(((size_t)e * S + s) * A + a) * (T * 2) + c];
    forecasts[((((size_t)env * H + ped) * k + j) * (T + 1) + 1) * 2 + c] = (double)v;
}

__global__ void fill_kernel(int B, int H, int T, int k, const uint8_t *__restrict__ in_cluster, const double *__restrict__ cv,
                            const double *__restrict__ cur, double uniform_logw, const double *__restrict__ logw_env,
                            double *__restrict__ forecasts, double *__restrict__ logw)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * H * k) return;
    const int j = (int)(i % k);
    const size_t eh = i / k;
    const int env = (int)(eh / H);
    double *f = forecasts + i * (size_t)(T + 1) * 2;
    f[0] = cur[eh * 2]; f[1] = cur[eh * 2 + 1];
    if (!in_cluster[eh])
        for (int t = 0; t < T * 2; ++t) f[2 + t] = cv[eh * T * 2 + t];
    logw[i] = logw_env ? logw_env[(size_t)env * k + j] : uniform_logw;
}

// ---------------------------------------------------------------------------------------------------------------
// KDE top-k, one CTA per environment of the bucket (fp32 like the reference).  For every horizon step t:
//   X [S,d] (d = 2A) -> cov = Xc^T Xc / (S-1); M = bw^-2 cov + 1e-6 I; P = inverse(M); L = cholesky(P) (lower);
//   Z = X L^-1 / bw;  ll_i = logsumexp_j(-0.5 |Z_i - Z_j|^2) - (0.5 d log 2pi + sum log diag L + log S); normalised over i;
// total = sum_t ll; keep the k largest in ascending order; logw = log-softmax of the kept totals.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KDE_THREADS = 128;

__device__ float block_max_sum_lse(const float *v, int n, float *scratch)
{ // logsumexp of v[0..n) computed by thread 0 and broadcast; n is tiny (<= S)
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = -INFINITY;
        for (int i = 0; i < n; ++i) m = fmaxf(m, v[i]);
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += expf(v[i] - m);
        scratch[0] = m + logf(s);
    }
    __syncthreads();
    return scratch[0];
}

__global__ void __launch_bounds__(KDE_THREADS)
kde_topk_kernel(const int32_t *__restrict__ order, int A, int S, int T, int k, const float *__restrict__ pos_b,
                int32_t *__restrict__ sel, double *__restrict__ logw_env, float *__restrict__ work, size_t work_per_env)
{
    const int e = blockIdx.x, env = order ? order[e] : e, tid = threadIdx.x, d = 2 * A;
    float *X = work + (size_t)e * work_per_env; // [S,d]
    float *M = X + S * d;                       // [d,d]   scaled covariance, then its Gauss-Jordan workspace
    float *P = M + d * d;                       // [d,d]   inverse
    float *L = P + d * d;                       // [d,d]   Cholesky factor, lower
    float *Li = L + d * d;                      // [d,d]   L^-1
    float *Z = Li + d * d;                      // [S,d]
    float *tot = Z + S * d;                     // [S]
    float *ll = tot + S;                        // [S]
    float *mean = ll + S;                       // [d]
    __shared__ float s_scr[4];
    __shared__ int s_piv;
    for (int i = tid; i < S; i += KDE_THREADS) tot[i] = 0.f;
    for (int t = 0; t < T; ++t) {
        const float bw = expf(logf(0.01f) + (T > 1 ? (float)t / (float)(T - 1) : 0.f) * (logf(0.1f) - logf(0.01f)));
        __syncthreads();
        for (int i = tid; i < S * d; i += KDE_THREADS) {
            const int s = i / d, c = i % d, a = c >> 1, xy = c & 1;
            X[i] = pos_b[((((size_t)e * S + s) * A + a) * T + t) * 2 + xy];
        }
        __syncthreads();
        for (int c = tid; c < d; c += KDE_THREADS) {
            float m = 0.f;
            for (int s = 0; s < S; ++s) m += X[s * d + c];
            mean[c] = m / (float)S;
        }
        __syncthreads();
        const float inv_bw2 = 1.0f / (bw * bw);
        for (int i = tid; i < d * d; i += KDE_THREADS) {
            const int r = i / d, c = i % d;
            float acc = 0.f;
            for (int s = 0; s < S; ++s) acc += (X[s * d + r] - mean[r]) * (X[s * d + c] - mean[c]);
            M[i] = inv_bw2 * (acc / (float)(S - 1)) + (r == c ? 1e-6f : 0.f);
            P[i] = r == c ? 1.f : 0.f;
        }
        __syncthreads();
        // Gauss-Jordan with partial pivoting: [M | P] -> [I | M^-1]
        for (int p = 0; p < d; ++p) {
            if (tid == 0) {
                int best = p; float bv = fabsf(M[p * d + p]);
                for (int r = p + 1; r < d; ++r) { const float v = fabsf(M[r * d + p]); if (v > bv) { bv = v; best = r; } }
                s_piv = best;
            }
            __syncthreads();
            const int pv = s_piv;
            if (pv != p)
                for (int c = tid; c < 2 * d; c += KDE_THREADS) {
                    float *Q = c < d ? M : P; const int cc = c < d ? c : c - d;
                    const float tmp = Q[p * d + cc]; Q[p * d + cc] = Q[pv * d + cc]; Q[pv * d + cc] = tmp;
                }
            __syncthreads();
            const float piv = M[p * d + p];
            __syncthreads();
            for (int c = tid; c < 2 * d; c += KDE_THREADS) {
                float *Q = c < d ? M : P; const int cc = c < d ? c : c - d;
                Q[p * d + cc] /= piv;
            }
            __syncthreads();
            for (int i = tid; i < d * 2 * d; i += KDE_THREADS) { // row r, column c of [M | P]
                const int r = i / (2 * d), c = i % (2 * d);
                if (r == p) continue;
                const float f = M[r * d + p];
                if (c == p && c < d) continue; // the pivot column is cleared after the sweep
                float *Q = c < d ? M : P; const int cc = c < d ? c : c - d;
                Q[r * d + cc] -= f * Q[p * d + cc];
            }
            __syncthreads();
            for (int r = tid; r < d; r += KDE_THREADS) if (r != p) M[r * d + p] = 0.f;
            __syncthreads();
        }
        // Cholesky of P (lower), column by column
        for (int i = tid; i < d * d; i += KDE_THREADS) { L[i] = 0.f; Li[i] = 0.f; }
        __syncthreads();
        for (int j = 0; j < d; ++j) {
            if (tid == 0) {
                float s = P[j * d + j];
                for (int q = 0; q < j; ++q) s -= L[j * d + q] * L[j * d + q];
                L[j * d + j] = sqrtf(s);
            }
            __syncthreads();
            const float djj = L[j * d + j];
            for (int r = j + 1 + tid; r < d; r += KDE_THREADS) {
                float s = P[r * d + j];
                for (int q = 0; q < j; ++q) s -= L[r * d + q] * L[j * d + q];
                L[r * d + j] = s / djj;
            }
            __syncthreads();
        }
        // L^-1 by forward substitution, one column per thread
        for (int c = tid; c < d; c += KDE_THREADS) {
            for (int r = c; r < d; ++r) {
                float s = r == c ? 1.f : 0.f;
                for (int q = c; q < r; ++q) s -= L[r * d + q] * Li[q * d + c];
                Li[r * d + c] = s / L[r * d + r];
            }
        }
        __syncthreads();
        for (int i = tid; i < S * d; i += KDE_THREADS) { // Z = X L^-1 / bw (row vector times matrix)
            const int s = i / d, c = i % d;
            float acc = 0.f;
            for (int q = c; q < d; ++q) acc += X[s * d + q] * Li[q * d + c];
            Z[i] = acc / bw;
        }
        if (tid == 0) {
            float ld = 0.f;
            for (int q = 0; q < d; ++q) ld += logf(L[q * d + q]);
            s_scr[1] = 0.5f * (float)d * logf(6.283185307179586f) + 0.5f * (2.0f * ld) + logf((float)S);
        }
        __syncthreads();
        const float Zc = s_scr[1];
        for (int i = tid; i < S; i += KDE_THREADS) {
            float m = -INFINITY;
            for (int j = 0; j < S; ++j) {
                float q = 0.f;
                for (int c = 0; c < d; ++c) { const float df = Z[i * d + c] - Z[j * d + c]; q += df * df; }
                m = fmaxf(m, -0.5f * q - Zc);
            }
            float s = 0.f;
            for (int j = 0; j < S; ++j) {
                float q = 0.f;
                for (int c = 0; c < d; ++c) { const float df = Z[i * d + c] - Z[j * d + c]; q += df * df; }
                s += expf(-0.5f * q - Zc - m);
            }
            ll[i] = m + logf(s);
        }
        const float norm = block_max_sum_lse(ll, S, s_scr);
        for (int i = tid; i < S; i += KDE_THREADS) tot[i] += ll[i] - norm;
        __syncthreads();
    }
    // top-k in ascending order of the total (torch.argsort(...)[-k:]); ties broken by index
    for (int i = tid; i < S; i += KDE_THREADS) {
        int rank = 0;
        for (int j = 0; j < S; ++j) rank += (tot[j] < tot[i]) || (tot[j] == tot[i] && j < i);
        if (rank >= S - k) { sel[(size_t)e * k + rank - (S - k)] = i; ll[rank - (S - k)] = tot[i]; }
    }
    const float norm = block_max_sum_lse(ll, k, s_scr);
    for (int j = tid; j < k; j += KDE_THREADS) logw_env[(size_t)env * k + j] = (double)(ll[j] - norm);
}

} // namespace

int snb_k_pred_noise(float *out, size_t n, uint64_t seed, uint64_t offset, cudaStream_t s)
{
    if (n == 0) return SNB_OK;
    const size_t q = (n + 3) / 4;
    noise_kernel<<<(unsigned)((q + 255) / 256), 256, 0, s>>>(out, n, seed, offset);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_gather(const int32_t *order, int cnt, int A, int H, int S, int T, const float *ctx, const float *noise, float *ctx_b,
                      float *xT_b, float *p0_b, const float *p0, cudaStream_t s)
{
    const size_t n = (size_t)cnt * (A * 256 + S * A * T * 2 + A * 2);
    if (n == 0) return SNB_OK;
    gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(order, cnt, A, H, S, T, ctx, noise, p0, ctx_b, xT_b, p0_b);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_scatter(const int32_t *order, int cnt, int A, int H, int S, int T, int k, const float *pos_b, const int32_t *sel,
                       const int32_t *ped_ids, double *forecasts, cudaStream_t s)
{
    const size_t n = (size_t)cnt * A * k * T * 2;
    if (n == 0) return SNB_OK;
    scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(order, cnt, A, H, S, T, k, pos_b, sel, ped_ids, forecasts);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_fill(int B, int H, int T, int k, const uint8_t *in_cluster, const double *cv, const double *cur, double uniform_logw,
                    const double *logw_env, double *forecasts, double *logw, cudaStream_t s)
{
    const size_t n = (size_t)B * H * k;
    if (n == 0) return SNB_OK;
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(B, H, T, k, in_cluster, cv, cur, uniform_logw, logw_env, forecasts, logw);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

size_t snb_k_pred_kde_work_floats(int cnt, int A, int S, int T)
{
    (void)T;
    const size_t d = 2 * (size_t)A;
    return (size_t)cnt * (2 * S * d + 4 * d * d + 2 * S + d);
}

int snb_k_pred_kde_topk(const int32_t *order, int cnt, int A, int S, int T, int k, const float *pos_b, int32_t *sel, double *logw_env,
                        float *work, cudaStream_t s)
{
    if (cnt == 0) return SNB_OK;
    const size_t d = 2 * (size_t)A;
    kde_topk_kernel<<<cnt, KDE_THREADS, 0, s>>>(order, A, S, T, k, pos_b, sel, logw_env, work, 2 * S * d + 4 * d * d + 2 * S + d);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
