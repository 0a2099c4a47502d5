// csrc/pred_encode.cu -- Trajectron context encoder of the JMID predictor in ONE kernel (fp32 SIMT):
//   node-history LSTM(6->128), PEDESTRIAN->PEDESTRIAN and PEDESTRIAN->JRDB_ROBOT edge LSTMs(12->128) on
//   [summed neighbour history | own history], edge mask, additive attention over the two edge types,
//   ctx = [combined edges | history]
// Reference: MultimodalGenerativeCVAE.obtain_encoded_tensors / encode_node_history / encode_edge /
// encode_total_edge_influence (MID/models/encoders/mgcvae.py:505-880), run_lstm_on_variable_length_seqs
// (model_utils.py:77-105, all histories are full here), AdditiveAttention (components/additive_attention.py:6-47).
//
// Tiling: one CTA = 32 rows (agents) x 512 threads; thread (u = tid & 127, rq = tid >> 7) owns hidden unit u of rows
// rq*8 .. rq*8+7: its four gate accumulators for 8 rows stay in registers, c stays in registers, h is exchanged through a
// double-buffered [128][32] shared tile read as broadcast float4.  Weights (transposed, 280 KB per LSTM) stream from L2.
#include "pred_internal.h"

namespace {

constexpr int HID = SNB_PRED_HID, TH = SNB_PRED_TH, ROWS = 32, RPT = 8; // rows per CTA, rows per thread

struct Smem {
    float xin[ROWS][TH][18];   // x_st | nb_ped | nb_rob
    float hbuf[2][HID][ROWS];
    float enc[3][HID][ROWS];   // 0: history, 1: ped edges (masked), 2: robot edges (masked)
    float emask[ROWS];
    float score[2][ROWS];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(512, 1)
pred_encode_kernel(PredEncDev W, const float *__restrict__ x_st, const float *__restrict__ nb_ped, const float *__restrict__ nb_rob,
                   const float *__restrict__ edge_mask, float *__restrict__ ctx, int rows)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, u = tid & (HID - 1), rq = tid >> 7;
    const int row0 = blockIdx.x * ROWS;
    // stage the inputs of the tile
    for (int i = tid; i < ROWS * TH * 18; i += 512) {
        const int r = i / (TH * 18), rem = i % (TH * 18), t = rem / 18, c = rem % 18;
        const int row = row0 + r;
        float v = 0.f;
        if (row < rows) {
            const size_t base = (size_t)row * (TH * 6) + t * 6;
            v = c < 6 ? x_st[base + c] : (c < 12 ? nb_ped[base + c - 6] : nb_rob[base + c - 12]);
        }
        sm.xin[r][t][c] = v;
    }
    if (tid < ROWS) sm.emask[tid] = (row0 + tid < rows) ? edge_mask[row0 + tid] : 0.f;
    __syncthreads();

    for (int l = 0; l < 3; ++l) {
        const PredLstmDev L = W.lstm[l];
        const float *wih = L.w_ihT, *whh = L.w_hhT;
        float bias[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) bias[g] = L.bias[g * HID + u];
        float c[RPT], hn[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) { c[r] = 0.f; hn[r] = 0.f; }
        int cur = 0;
        for (int t = 0; t < TH; ++t) {
            float acc[4][RPT];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int r = 0; r < RPT; ++r) acc[g][r] = bias[g];
            // input part: edge encoders see [neighbour sum (6) | own history (6)]
            for (int k = 0; k < L.din; ++k) {
                const int col = l == 0 ? k : (k < 6 ? (l == 1 ? 6 + k : 12 + k) : k - 6);
                float wv[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) wv[g] = __ldg(wih + (size_t)k * 512 + g * HID + u);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const float xv = sm.xin[rq * RPT + r][t][col];
#pragma unroll
                    for (int g = 0; g < 4; ++g) acc[g][r] = fmaf(wv[g], xv, acc[g][r]);
                }
            }
            if (t > 0) { // h_0 = 0
#pragma unroll 4
                for (int k = 0; k < HID; ++k) {
                    float wv[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) wv[g] = __ldg(whh + (size_t)k * 512 + g * HID + u);
                    const float4 h0 = *reinterpret_cast<const float4 *>(&sm.hbuf[cur][k][rq * RPT]);
                    const float4 h1 = *reinterpret_cast<const float4 *>(&sm.hbuf[cur][k][rq * RPT + 4]);
                    const float hv[RPT] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                    for (int r = 0; r < RPT; ++r)
#pragma unroll
                        for (int g = 0; g < 4; ++g) acc[g][r] = fmaf(wv[g], hv[r], acc[g][r]);
                }
            }
            const int nxt = cur ^ 1;
#pragma unroll
            for (int r = 0; r < RPT; ++r) { // torch gate order i, f, g, o
                c[r] = sigmoidf_(acc[1][r]) * c[r] + sigmoidf_(acc[0][r]) * tanhf(acc[2][r]);
                hn[r] = sigmoidf_(acc[3][r]) * tanhf(c[r]);
                sm.hbuf[nxt][u][rq * RPT + r] = hn[r];
            }
            cur = nxt;
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) sm.enc[l][u][rq * RPT + r] = l == 0 ? hn[r] : hn[r] * sm.emask[rq * RPT + r];
        __syncthreads();
    }

    // additive attention: score_e = v . tanh(W1 enc_e + W2 hist), softmax over the two edge types
    float q2[RPT], s1[RPT], s2[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) { q2[r] = 0.f; s1[r] = 0.f; s2[r] = 0.f; }
    for (int k = 0; k < HID; ++k) {
        const float w1 = __ldg(W.w1T + (size_t)k * HID + u), w2 = __ldg(W.w2T + (size_t)k * HID + u);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            q2[r] = fmaf(w2, sm.enc[0][k][rq * RPT + r], q2[r]);
            s1[r] = fmaf(w1, sm.enc[1][k][rq * RPT + r], s1[r]);
            s2[r] = fmaf(w1, sm.enc[2][k][rq * RPT + r], s2[r]);
        }
    }
    const float vu = __ldg(W.v + u);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        sm.hbuf[0][u][rq * RPT + r] = vu * tanhf(s1[r] + q2[r]);
        sm.hbuf[1][u][rq * RPT + r] = vu * tanhf(s2[r] + q2[r]);
    }
    __syncthreads();
    if (tid < 2 * ROWS) {
        const int e = tid >> 5, r = tid & 31;
        float acc = 0.f;
        for (int k = 0; k < HID; ++k) acc += sm.hbuf[e][k][r];
        sm.score[e][r] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int rr = rq * RPT + r, row = row0 + rr;
        if (row >= rows) continue;
        const float a = sm.score[0][rr], b = sm.score[1][rr], m = fmaxf(a, b);
        const float ea = expf(a - m), eb = expf(b - m), inv = 1.0f / (ea + eb);
        ctx[(size_t)row * 256 + u] = (ea * inv) * sm.enc[1][u][rr] + (eb * inv) * sm.enc[2][u][rr];
        ctx[(size_t)row * 256 + HID + u] = sm.enc[0][u][rr];
    }
}

__global__ void transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    dst[(size_t)c * rows + r] = src[i];
}

__global__ void add_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = a[i] + b[i];
}

} // namespace

int snb_k_pred_encode(const PredEncDev *w, const float *x_st, const float *nb_ped, const float *nb_rob, const float *edge_mask,
                      float *ctx, int rows, cudaStream_t s)
{
    if (rows == 0) return SNB_OK;
    static bool attr_set = false;
    if (!attr_set) {
        SNB_CUDA_TRY(cudaFuncSetAttribute(pred_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        attr_set = true;
    }
    pred_encode_kernel<<<(rows + ROWS - 1) / ROWS, 512, sizeof(Smem), s>>>(*w, x_st, nb_ped, nb_rob, edge_mask, ctx, rows);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_transpose_f32(const float *src, float *dst, int rows, int cols, cudaStream_t s)
{
    const int n = rows * cols;
    transpose_kernel<<<(n + 255) / 256, 256, 0, s>>>(src, dst, rows, cols);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_add_f32(const float *a, const float *b, float *dst, int n, cudaStream_t s)
{
    add_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, dst, n);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
