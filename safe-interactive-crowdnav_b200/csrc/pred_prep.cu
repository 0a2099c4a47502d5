// csrc/pred_prep.cu -- JMID predictor pre-processing on the device, one warp per environment (fp64, compiled with
// -fmad=false so that every threshold decision is taken on the same IEEE values as numpy takes it):
//   history ring push                      ForecasterSimSuper.update_state_hists      mid_sim_wrapper.py:172-204
//   finite-difference node states          derivative_of                              MID/environment/data_utils.py:24-37
//   3 m clustering, robot-nearest cluster  convert_to_mid_state_env                   mid_sim_wrapper.py:322-355
//   constant-velocity fall-back            convert_to_mid_state_env                   mid_sim_wrapper.py:413-429
//   scene graph + edge scaling             TemporalSceneGraph.*                       MID/environment/scene_graph.py:111-225
//   standardised node / neighbour tensors  get_node_timestep_data                     MID/dataset/preprocessing.py:428-620
// Lane 0 of a warp is the robot (track id -1 sorts first), lane i >= 1 is human i-1; H <= 31.
#include "pred_internal.h"

namespace {

constexpr int TH = SNB_PRED_TH;       // 6 history frames
constexpr int ENVS_PER_CTA = 4;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// states [6 frames][6] = x, y, vx, vy, ax, ay
__device__ __forceinline__ void node_states(const double *px, const double *py, double dt, double st[TH][6])
{
    double vx[TH], vy[TH];
    for (int t = 0; t < TH; ++t) {
        const int a = t == 0 ? 1 : t, b = t == 0 ? 0 : t - 1; // ediff1d(to_begin = x[1] - x[0])
        vx[t] = (px[a] - px[b]) / dt;
        vy[t] = (py[a] - py[b]) / dt;
    }
    for (int t = 0; t < TH; ++t) {
        const int a = t == 0 ? 1 : t, b = t == 0 ? 0 : t - 1;
        st[t][0] = px[t]; st[t][1] = py[t]; st[t][2] = vx[t]; st[t][3] = vy[t];
        st[t][4] = (vx[a] - vx[b]) / dt;
        st[t][5] = (vy[a] - vy[b]) / dt;
    }
}

__global__ void __launch_bounds__(32 * ENVS_PER_CTA)
pred_prep_kernel(const double *__restrict__ hist, const double *__restrict__ robot_hist, int B, int H, double radius, double pos_std, double dt,
                 int horizon, PredPrepOut o)
{
    __shared__ double s_st[ENVS_PER_CTA][32][TH][6];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * ENVS_PER_CTA + w;
    if (env >= B) return;
    const int n = H + 1;
    const bool live = lane < n;
    double px[TH], py[TH];
    for (int t = 0; t < TH; ++t) { px[t] = 0.0; py[t] = 0.0; }
    if (live) {
        const double *src = lane == 0 ? robot_hist + (size_t)env * TH * 2 : hist + ((size_t)env * H + (lane - 1)) * TH * 2;
        for (int t = 0; t < TH; ++t) { px[t] = src[2 * t]; py[t] = src[2 * t + 1]; }
    }
    const double x = px[TH - 1], y = py[TH - 1];
    // ---- pairwise distance mask on the last frame (strict <), cluster means, cluster nearest the robot ----
    unsigned row = 0u;
    double sx = 0.0, sy = 0.0;
    int cnt = 0;
    for (int j = 0; j < n; ++j) {
        const double xj = shfl_d(x, j), yj = shfl_d(y, j);
        const double dx = x - xj, dy = y - yj;
        const double d = sqrt(dx * dx + dy * dy);
        if (d < radius) { row |= 1u << j; sx += xj; sy += yj; ++cnt; }
    }
    const double x0 = shfl_d(x, 0), y0 = shfl_d(y, 0);
    double rd = 1.0e300;
    if (live && lane >= 1) {
        const double mx = sx / (double)cnt - x0, my = sy / (double)cnt - y0;
        rd = sqrt(mx * mx + my * my);
    }
    int best = lane;
    for (int off = 16; off > 0; off >>= 1) { // argmin, first index on ties (np.argmin)
        const double r2 = shfl_d(rd, lane ^ off);
        const int b2 = __shfl_sync(0xffffffffu, best, lane ^ off);
        if (r2 < rd || (r2 == rd && b2 < best)) { rd = r2; best = b2; }
    }
    const unsigned in_mask = __shfl_sync(0xffffffffu, row, best);
    const bool in_cluster = live && ((in_mask >> lane) & 1u);
    // ---- node states of every node into shared memory ----
    double st[TH][6];
    node_states(px, py, dt, st);
    for (int t = 0; t < TH; ++t)
        for (int c = 0; c < 6; ++c) s_st[w][lane][t][c] = st[t][c];
    __syncwarp();
    const unsigned ped_mask = in_mask & ~1u;
    const int A = __popc(ped_mask);
    const size_t slot0 = (size_t)env * H;
    if (lane == 0) o.n_in[env] = A;
    if (live && lane >= 1) {
        const int h = lane - 1;
        o.in_cluster[slot0 + h] = in_cluster ? 1 : 0;
        // constant-velocity forecast and the current pose of EVERY human (rows of the humans outside the cluster are used)
        const double vdx = st[TH - 1][2] * dt, vdy = st[TH - 1][3] * dt;
        double cx = 0.0, cy = 0.0;
        for (int t = 0; t < horizon; ++t) {
            cx = t == 0 ? vdx : cx + vdx; // np.cumsum of a constant vector
            cy = t == 0 ? vdy : cy + vdy;
            o.cv[((slot0 + h) * horizon + t) * 2 + 0] = x + cx;
            o.cv[((slot0 + h) * horizon + t) * 2 + 1] = y + cy;
        }
        o.cur[(slot0 + h) * 2 + 0] = x;
        o.cur[(slot0 + h) * 2 + 1] = y;
    }
    const double stdv[6] = {pos_std, pos_std, 2.0, 2.0, 1.0, 1.0}; // position std := the attention radius (preprocessing.py:477-478, 540)
    if (in_cluster && lane >= 1) {
        const int slot = __popc(ped_mask & ((1u << lane) - 1u));
        const size_t row_i = slot0 + slot;
        double nbp[TH][6], nbr[TH][6];
        for (int t = 0; t < TH; ++t)
            for (int c = 0; c < 6; ++c) { nbp[t][c] = 0.0; nbr[t][c] = 0.0; }
        double esum = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == lane || !((in_mask >> j) & 1u)) continue;
            // adjacency (<= radius) on frames t-2, t-1, t; edge-addition filter [.25,.5,.75,1], removal filter [1,0]
            double wsum = 0.0;
            bool cur_adj = false;
            for (int k = 0; k < 3; ++k) {
                const int t = TH - 3 + k;
                const double dx = st[t][0] - s_st[w][j][t][0], dy = st[t][1] - s_st[w][j][t][1];
                const bool adj = sqrt(dx * dx + dy * dy) <= radius;
                if (adj) wsum += k == 2 ? 0.25 : (k == 1 ? 0.5 : 0.75);
                if (k == 2) cur_adj = adj;
            }
            double es = wsum < 1.0 ? wsum : 1.0;
            if (!cur_adj) es = 0.0;
            if (!(es > 1e-2)) continue;
            esum += es;
            for (int t = 0; t < TH; ++t)
                for (int c = 0; c < 6; ++c) {
                    const double v = (s_st[w][j][t][c] - st[TH - 1][c]) / stdv[c];
                    if (j == 0) nbr[t][c] += v; else nbp[t][c] += v;
                }
        }
        float *xs = o.x_st + row_i * (TH * 6), *np_ = o.nb_ped + row_i * (TH * 6), *nr = o.nb_rob + row_i * (TH * 6);
        for (int t = 0; t < TH; ++t)
            for (int c = 0; c < 6; ++c) {
                const double rel = c == 0 ? x : (c == 1 ? y : 0.0);
                xs[t * 6 + c] = (float)((st[t][c] - rel) / stdv[c]);
                np_[t * 6 + c] = (float)nbp[t][c];
                nr[t * 6 + c] = (float)nbr[t][c];
            }
        o.edge_mask[row_i] = (float)(esum < 1.0 ? esum : 1.0);
        o.p0[row_i * 2 + 0] = (float)x;
        o.p0[row_i * 2 + 1] = (float)y;
        o.ped_ids[row_i] = lane - 1;
    }
    // unused slots: zero inputs so that the encoder tile never reads garbage
    if (lane >= A && lane < H) {
        const size_t row_i = slot0 + lane;
        for (int k = 0; k < TH * 6; ++k) { o.x_st[row_i * (TH * 6) + k] = 0.f; o.nb_ped[row_i * (TH * 6) + k] = 0.f; o.nb_rob[row_i * (TH * 6) + k] = 0.f; }
        o.edge_mask[row_i] = 0.f;
        o.p0[row_i * 2] = 0.f; o.p0[row_i * 2 + 1] = 0.f;
        o.ped_ids[row_i] = -1;
    }
}

// ring push: frames shift one to the left, the new positions land in frame TH-1 (histories are capped at 6 frames,
// mid_sim_wrapper.py:199-202); a fresh ring (count == 0) is filled with the first observation.
__global__ void pred_push_kernel(double *__restrict__ hist, double *__restrict__ robot_hist, const double *__restrict__ hpx,
                                 const double *__restrict__ hpy, const double *__restrict__ rpx, const double *__restrict__ rpy,
                                 int B, int H, int first)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (H + 1)) return;
    const int env = i / (H + 1), k = i % (H + 1);
    double *dst = k == 0 ? robot_hist + (size_t)env * TH * 2 : hist + ((size_t)env * H + (k - 1)) * TH * 2;
    const double nx = k == 0 ? rpx[env] : hpx[(size_t)env * H + (k - 1)];
    const double ny = k == 0 ? rpy[env] : hpy[(size_t)env * H + (k - 1)];
    for (int t = 0; t < TH - 1; ++t) {
        dst[2 * t] = first ? nx : dst[2 * (t + 1)];
        dst[2 * t + 1] = first ? ny : dst[2 * (t + 1) + 1];
    }
    dst[2 * (TH - 1)] = nx;
    dst[2 * (TH - 1) + 1] = ny;
}

// numpy's pairwise summation of a strided 1-D double vector (n <= 128 -> eight running sums), so that np.mean is reproduced
__device__ double np_pairwise_sum(const double *a, int n, size_t stride)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i * stride];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i * stride];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2, stride) + np_pairwise_sum(a + n2 * stride, n - n2, stride);
}

// SICNavAcados.predict ingest (sicnav_acados.py:1645-1667): one thread per (env, human)
__global__ void pred_ingest_kernel(const double *__restrict__ forecasts, const double *__restrict__ logw, int B, int H, int k, int T,
                                   int horiz, double dt, int joint, double *__restrict__ resh, double *__restrict__ weights,
                                   double *__restrict__ goals, double *__restrict__ vpref)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int env = i / H, h = i % H;
    const int Tp = T < horiz + 1 ? T : horiz + 1;
    const double *f = forecasts + (size_t)i * k * (T + 1) * 2; // [k, T+1, 2], frame 0 = current pose (dropped, :1645)
    for (int t = 0; t < Tp; ++t)
        for (int j = 0; j < k; ++j)
            for (int c = 0; c < 2; ++c)
                resh[(((size_t)env * Tp + t) * (H * k) + (size_t)h * k + j) * 2 + c] = f[((size_t)j * (T + 1) + 1 + t) * 2 + c];
    const size_t stride = (size_t)(T + 1) * 2;
    goals[(size_t)i * 2 + 0] = np_pairwise_sum(f + 2, k, stride) / (double)k;
    goals[(size_t)i * 2 + 1] = np_pairwise_sum(f + 3, k, stride) / (double)k;
    double vm = 0.0;
    for (int j = 0; j < k; ++j)
        for (int t = 1; t < T; ++t) {
            const double dx = f[((size_t)j * (T + 1) + t + 1) * 2] - f[((size_t)j * (T + 1) + t) * 2];
            const double dy = f[((size_t)j * (T + 1) + t + 1) * 2 + 1] - f[((size_t)j * (T + 1) + t) * 2 + 1];
            const double v = sqrt(dx * dx + dy * dy) / dt;
            vm = (j == 0 && t == 1) ? v : (v > vm ? v : vm);
        }
    vpref[i] = vm;
    if (joint) {
        if (h == 0)
            for (int j = 0; j < k; ++j) weights[(size_t)env * k + j] = logw[(size_t)i * k + j];
    } else {
        for (int j = 0; j < k; ++j) weights[(size_t)i * k + j] = logw[(size_t)i * k + j];
    }
}

// convert_to_mpc_state_vector (sicnav_acados.py:222-289) on the joint state SICNavAcados.predict builds (:1655-1681): one thread per env
__global__ void pred_mpc_state_kernel(const double *__restrict__ robot, const double *__restrict__ humans, const double *__restrict__ goals,
                                      const double *__restrict__ weights, int B, int H, int k, int joint, double *__restrict__ out,
                                      double *__restrict__ theta)
{
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= B) return;
    const int nx_hum = joint ? 6 : 6 + k, nx = 10 + nx_hum * H + (joint ? k : 0);
    const double *r = robot + (size_t)env * 9; // px py theta lvel omega v_dot omega_dot gx gy
    double *v = out + (size_t)env * nx;
    v[0] = r[0]; v[1] = r[1]; v[2] = sin(r[2]); v[3] = cos(r[2]);
    v[4] = r[3]; v[5] = r[4]; v[6] = r[5]; v[7] = r[6]; v[8] = r[7]; v[9] = r[8];
    for (int h = 0; h < H; ++h) {
        const double *hs = humans + ((size_t)env * H + h) * 4; // px py vx vy
        double *o = v + 10 + (size_t)h * nx_hum;
        o[0] = hs[0]; o[1] = hs[1]; o[2] = hs[2]; o[3] = hs[3];
        o[4] = goals[((size_t)env * H + h) * 2]; o[5] = goals[((size_t)env * H + h) * 2 + 1];
        if (!joint)
            for (int j = 0; j < k; ++j) o[6 + j] = weights[((size_t)env * H + h) * k + j];
        if (theta) theta[(size_t)env * H + h] = (hs[2] != 0.0 || hs[3] != 0.0) ? atan2(hs[3], hs[2]) : 0.0; // :1678
    }
    if (joint)
        for (int j = 0; j < k; ++j) v[nx - k + j] = weights[(size_t)env * k + j];
}

// the per-stage parameter vector handed to the solver (sicnav_acados.py:1389-1413): one thread per (env, stage, column)
__global__ void pred_stage_params_kernel(const double *__restrict__ resh, const double *__restrict__ prefix, const double *__restrict__ stat,
                                         int B, int horiz, int Tp, int HK, int n_prefix, int n_stat, double *__restrict__ out)
{
    const int np = n_prefix + 4 * HK + n_stat;
    const size_t total = (size_t)B * (horiz + 1) * np;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % np), st = (int)((i / np) % (horiz + 1)), env = (int)(i / ((size_t)np * (horiz + 1)));
        double val;
        if (c < n_prefix) {
            val = prefix[((size_t)env * (horiz + 1) + st) * n_prefix + c];
        } else if (c < n_prefix + 4 * HK) {
            const int q = (c - n_prefix) / HK, j = (c - n_prefix) % HK;          // q: X_t[:,0], X_t[:,1], X_t+1[:,0], X_t+1[:,1]
            const int t = (st < horiz ? st : horiz - 1) + (q >> 1);               // the terminal stage reuses idx = horiz - 1 (:1403-1405)
            val = resh[(((size_t)env * Tp + t) * HK + j) * 2 + (q & 1)];
        } else {
            val = stat[c - n_prefix - 4 * HK];
        }
        out[i] = val;
    }
}

// CrowdSimPlus.step's `self.states.append([...])` (crowd_sim_plus.py:1175-1181) as a ring of L frames: slot <- positions before the step
__global__ void state_log_push_kernel(const double *__restrict__ hpx, const double *__restrict__ hpy, const double *__restrict__ rpx,
                                      const double *__restrict__ rpy, int B, int H, int L, int slot, double *__restrict__ log)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (H + 1)) return;
    const int env = i / (H + 1), a = i % (H + 1);
    double *o = log + (((size_t)env * L + slot) * (H + 1) + a) * 2;
    if (a < H) { o[0] = hpx[(size_t)env * H + a]; o[1] = hpy[(size_t)env * H + a]; }
    else { o[0] = rpx[env]; o[1] = rpy[env]; }
}

// reset_scenario_values' history bootstrap (sicnav_acados.py:1163-1182): states[-Th-1:-1] -> the six ring frames, oldest first
__global__ void pred_bootstrap_kernel(const double *__restrict__ log, int B, int H, int L, int newest, double *__restrict__ hist,
                                      double *__restrict__ robot_hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (H + 1) * TH) return;
    const int f = i % TH, a = (i / TH) % (H + 1), env = i / (TH * (H + 1));
    const int slot = ((newest - TH + f) % L + L) % L;        // frame f of states[-Th-1:-1]: the newest entry (states[-1]) is skipped
    const double *src = log + (((size_t)env * L + slot) * (H + 1) + a) * 2;
    double *dst = a < H ? hist + (((size_t)env * H + a) * TH + f) * 2 : robot_hist + ((size_t)env * TH + f) * 2;
    dst[0] = src[0]; dst[1] = src[1];
}

} // namespace

int snb_k_pred_mpc_state(const double *robot, const double *humans, const double *goals, const double *weights, int B, int H, int k,
                         int joint, double *out, double *theta, cudaStream_t s)
{
    if (B == 0) return SNB_OK;
    pred_mpc_state_kernel<<<(B + 127) / 128, 128, 0, s>>>(robot, humans, goals, weights, B, H, k, joint, out, theta);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_stage_params(const double *resh, const double *prefix, const double *stat, int B, int horiz, int Tp, int HK, int n_prefix,
                            int n_stat, double *out, cudaStream_t s)
{
    const size_t total = (size_t)B * (horiz + 1) * (n_prefix + 4 * HK + n_stat);
    if (total == 0) return SNB_OK;
    const size_t blocks = (total + 255) / 256;
    pred_stage_params_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, s>>>(resh, prefix, stat, B, horiz, Tp, HK, n_prefix, n_stat, out);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_state_log_push(const double *hpx, const double *hpy, const double *rpx, const double *rpy, int B, int H, int L, int slot,
                         double *log, cudaStream_t s)
{
    if (B == 0) return SNB_OK;
    const int n = B * (H + 1);
    state_log_push_kernel<<<(n + 127) / 128, 128, 0, s>>>(hpx, hpy, rpx, rpy, B, H, L, slot, log);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_bootstrap(const double *log, int B, int H, int L, int newest, double *hist, double *robot_hist, cudaStream_t s)
{
    if (B == 0) return SNB_OK;
    const int n = B * (H + 1) * TH;
    pred_bootstrap_kernel<<<(n + 127) / 128, 128, 0, s>>>(log, B, H, L, newest, hist, robot_hist);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_prep(const double *hist, const double *robot_hist, int B, int H, double radius, double pos_std, double dt, int horizon,
                    const PredPrepOut *out, cudaStream_t s)
{
    if (B == 0) return SNB_OK;
    pred_prep_kernel<<<(B + ENVS_PER_CTA - 1) / ENVS_PER_CTA, 32 * ENVS_PER_CTA, 0, s>>>(hist, robot_hist, B, H, radius, pos_std, dt, horizon, *out);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_push(double *hist, double *robot_hist, const double *hpx, const double *hpy, const double *rpx, const double *rpy,
                    int B, int H, int first, cudaStream_t s)
{
    if (B == 0) return SNB_OK;
    const int n = B * (H + 1);
    pred_push_kernel<<<(n + 127) / 128, 128, 0, s>>>(hist, robot_hist, hpx, hpy, rpx, rpy, B, H, first);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

int snb_k_pred_ingest(const double *forecasts, const double *logw, int B, int H, int k, int T, int horiz, double dt, int joint,
                      double *resh, double *weights, double *goals, double *vpref, cudaStream_t s)
{
    if (B == 0) return SNB_OK;
    const int n = B * H;
    pred_ingest_kernel<<<(n + 127) / 128, 128, 0, s>>>(forecasts, logw, B, H, k, T, horiz, dt, joint, resh, weights, goals, vpref);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
