// csrc/pred_api.cu -- C-ABI entry points of the JMID predictor around the denoiser (include/snb.h): history rings,
// pre-processing + context encoder, the batched predict_ret_best and the MPC ingest.
#include <cmath>
#include <cstring>
#include <vector>

#include "pred_internal.h"

struct SnbPredictor {
    int max_envs = 0, H = 0, S = 0, T = 0, joint = 1;
    SnbJmid *den = nullptr;
    std::vector<void *> allocs;
    PredEncDev enc;
    // history rings
    double *hist = nullptr, *robot_hist = nullptr;
    long n_pushed = 0;
    // pre-processing outputs and bucket staging
    PredPrepOut prep;
    float *ctx = nullptr, *noise = nullptr, *ctx_b = nullptr, *xT_b = nullptr, *vel_b = nullptr, *pos_b = nullptr, *p0_b = nullptr;
    float *kde_work = nullptr;
    size_t kde_work_floats = 0;
    int32_t *order = nullptr, *sel = nullptr;
    double *logw_env = nullptr;
    int32_t *h_n_in = nullptr, *h_order = nullptr; // pinned
    uint64_t calls = 0;
    double pos_std = 0.0; // 0: the attention radius, as the reference (preprocessing.py:477-478)
    cudaStream_t own_stream = nullptr;
    double *d_fc = nullptr, *d_lw = nullptr; // host-call staging
    size_t d_fc_elems = 0;
};

namespace {

template <class Tp>
int palloc(SnbPredictor *p, Tp **out, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(Tp));
    if (e != cudaSuccess) { snb_set_error("snb_pred: cudaMalloc(%zu B) failed: %s", n * sizeof(Tp), cudaGetErrorString(e)); return SNB_ENOMEM; }
    p->allocs.push_back(q);
    *out = static_cast<Tp *>(q);
    return SNB_OK;
}

int make_lstm(SnbPredictor *p, PredLstmDev *dst, const SnbLstmWeights *w, int din, cudaStream_t s)
{
    SNB_REQUIRE(w->w_ih && w->w_hh && w->b_ih && w->b_hh, SNB_EINVAL, "snb_pred_create: NULL LSTM weight pointer");
    float *ihT = nullptr, *hhT = nullptr, *bias = nullptr;
    int rc = palloc(p, &ihT, (size_t)din * 512);
    if (!rc) rc = palloc(p, &hhT, (size_t)128 * 512);
    if (!rc) rc = palloc(p, &bias, 512);
    if (!rc) rc = snb_k_transpose_f32(w->w_ih, ihT, 512, din, s);
    if (!rc) rc = snb_k_transpose_f32(w->w_hh, hhT, 512, 128, s);
    if (!rc) rc = snb_k_add_f32(w->b_ih, w->b_hh, bias, 512, s);
    dst->w_ihT = ihT; dst->w_hhT = hhT; dst->bias = bias; dst->din = din;
    return rc;
}

int run_encode(SnbPredictor *p, int B, double radius, double dt, cudaStream_t s)
{
    int rc = snb_k_pred_prep(p->hist, p->robot_hist, B, p->H, radius, p->pos_std > 0.0 ? p->pos_std : radius, dt, p->T, &p->prep, s);
    if (rc) return rc;
    return snb_k_pred_encode(&p->enc, p->prep.x_st, p->prep.nb_ped, p->prep.nb_rob, p->prep.edge_mask, p->ctx, B * p->H, s);
}

} // namespace

extern "C" int snb_pred_create(SnbPredictor **out, const SnbEncoderWeights *w, SnbJmid *denoiser, int32_t max_envs, int32_t H, void *stream)
{
    SNB_REQUIRE(out && w && denoiser, SNB_EINVAL, "snb_pred_create: NULL argument");
    int32_t A = 0, S = 0, T = 0, joint = 0;
    int rc = snb_jmid_dims(denoiser, &A, &S, &T, &joint);
    if (rc) return rc;
    SNB_REQUIRE(max_envs >= 1 && H >= 1 && H <= 31, SNB_EINVAL, "snb_pred_create: need max_envs >= 1 and 1 <= H <= 31 (one warp per environment)");
    SNB_REQUIRE(H <= A, SNB_EINVAL, "snb_pred_create: H=%d exceeds the denoiser's agent capacity A=%d", H, A);
    SNB_REQUIRE(w->att_w1 && w->att_w2 && w->att_v, SNB_EINVAL, "snb_pred_create: NULL attention weight pointer");
    cudaStream_t s = (cudaStream_t)stream;
    SnbPredictor *p = new SnbPredictor();
    p->max_envs = max_envs; p->H = H; p->S = S; p->T = T; p->joint = joint; p->den = denoiser;
#define TRY(x) do { if (!rc) rc = (x); } while (0)
    TRY(make_lstm(p, &p->enc.lstm[0], &w->node_history, 6, s));
    TRY(make_lstm(p, &p->enc.lstm[1], &w->edge_ped, 12, s));
    TRY(make_lstm(p, &p->enc.lstm[2], &w->edge_robot, 12, s));
    float *w1T = nullptr, *w2T = nullptr, *v = nullptr;
    TRY(palloc(p, &w1T, 128 * 128)); TRY(palloc(p, &w2T, 128 * 128)); TRY(palloc(p, &v, 128));
    TRY(snb_k_transpose_f32(w->att_w1, w1T, 128, 128, s));
    TRY(snb_k_transpose_f32(w->att_w2, w2T, 128, 128, s));
    if (!rc && cudaMemcpyAsync(v, w->att_v, 128 * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = SNB_ECUDA;
    p->enc.w1T = w1T; p->enc.w2T = w2T; p->enc.v = v;
    const size_t BH = (size_t)max_envs * H, traj = BH * S * T * 2;
    TRY(palloc(p, &p->hist, BH * SNB_PRED_TH * 2));
    TRY(palloc(p, &p->robot_hist, (size_t)max_envs * SNB_PRED_TH * 2));
    TRY(palloc(p, &p->prep.n_in, (size_t)max_envs));
    TRY(palloc(p, &p->prep.in_cluster, BH));
    TRY(palloc(p, &p->prep.ped_ids, BH));
    TRY(palloc(p, &p->prep.x_st, BH * 36)); TRY(palloc(p, &p->prep.nb_ped, BH * 36)); TRY(palloc(p, &p->prep.nb_rob, BH * 36));
    TRY(palloc(p, &p->prep.edge_mask, BH));
    TRY(palloc(p, &p->prep.p0, BH * 2));
    TRY(palloc(p, &p->prep.cv, BH * T * 2));
    TRY(palloc(p, &p->prep.cur, BH * 2));
    TRY(palloc(p, &p->ctx, BH * 256));
    TRY(palloc(p, &p->noise, traj));
    TRY(palloc(p, &p->ctx_b, BH * 256));
    TRY(palloc(p, &p->xT_b, traj)); TRY(palloc(p, &p->vel_b, traj)); TRY(palloc(p, &p->pos_b, traj));
    TRY(palloc(p, &p->p0_b, BH * 2));
    TRY(palloc(p, &p->order, (size_t)max_envs));
    TRY(palloc(p, &p->sel, (size_t)max_envs * S));
    TRY(palloc(p, &p->logw_env, (size_t)max_envs * S));
    p->kde_work_floats = snb_k_pred_kde_work_floats(max_envs, H, S, T);
    TRY(palloc(p, &p->kde_work, p->kde_work_floats));
#undef TRY
    if (!rc && cudaMallocHost(&p->h_n_in, sizeof(int32_t) * max_envs) != cudaSuccess) rc = SNB_ENOMEM;
    if (!rc && cudaMallocHost(&p->h_order, sizeof(int32_t) * max_envs) != cudaSuccess) rc = SNB_ENOMEM;
    if (!rc && cudaStreamSynchronize(s) != cudaSuccess) { snb_set_error("snb_pred_create: %s", cudaGetErrorString(cudaGetLastError())); rc = SNB_ECUDA; }
    if (rc) { snb_pred_destroy(p); return rc; }
    *out = p;
    return SNB_OK;
}

extern "C" int snb_pred_destroy(SnbPredictor *p)
{
    if (!p) return SNB_OK;
    for (void *q : p->allocs) cudaFree(q);
    if (p->h_n_in) cudaFreeHost(p->h_n_in);
    if (p->h_order) cudaFreeHost(p->h_order);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    cudaFree(p->d_fc); cudaFree(p->d_lw);
    delete p;
    return SNB_OK;
}

extern "C" int snb_pred_set_position_std(SnbPredictor *p, double pos_std)
{
    SNB_REQUIRE(p && pos_std >= 0.0, SNB_EINVAL, "snb_pred_set_position_std: bad argument");
    p->pos_std = pos_std;
    return SNB_OK;
}

extern "C" int snb_pred_push_history(SnbPredictor *p, const double *hpx, const double *hpy, const double *rpx, const double *rpy, int32_t B,
                                     void *stream)
{
    SNB_REQUIRE(p && hpx && hpy && rpx && rpy, SNB_EINVAL, "snb_pred_push_history: NULL argument");
    SNB_REQUIRE(B >= 0 && B <= p->max_envs, SNB_EINVAL, "snb_pred_push_history: B=%d beyond max_envs=%d", B, p->max_envs);
    int rc = snb_k_pred_push(p->hist, p->robot_hist, hpx, hpy, rpx, rpy, B, p->H, p->n_pushed == 0, (cudaStream_t)stream);
    if (!rc) ++p->n_pushed;
    return rc;
}

extern "C" int snb_pred_reset_history(SnbPredictor *p)
{
    SNB_REQUIRE(p, SNB_EINVAL, "snb_pred_reset_history: NULL handle");
    p->n_pushed = 0;
    return SNB_OK;
}

extern "C" int snb_pred_set_history(SnbPredictor *p, const double *hist, const double *robot_hist, int32_t B, void *stream)
{
    SNB_REQUIRE(p && hist && robot_hist, SNB_EINVAL, "snb_pred_set_history: NULL argument");
    SNB_REQUIRE(B >= 0 && B <= p->max_envs, SNB_EINVAL, "snb_pred_set_history: B=%d beyond max_envs=%d", B, p->max_envs);
    cudaStream_t s = (cudaStream_t)stream;
    SNB_CUDA_TRY(cudaMemcpyAsync(p->hist, hist, (size_t)B * p->H * SNB_PRED_TH * 2 * sizeof(double), cudaMemcpyDefault, s));
    SNB_CUDA_TRY(cudaMemcpyAsync(p->robot_hist, robot_hist, (size_t)B * SNB_PRED_TH * 2 * sizeof(double), cudaMemcpyDefault, s));
    p->n_pushed = SNB_PRED_TH;
    return SNB_OK;
}

extern "C" int snb_pred_encode(SnbPredictor *p, int32_t B, double radius, double dt, float *ctx, int32_t *n_in, int32_t *ped_ids,
                               uint8_t *in_cluster, void *stream)
{
    SNB_REQUIRE(p, SNB_EINVAL, "snb_pred_encode: NULL handle");
    SNB_REQUIRE(B >= 0 && B <= p->max_envs, SNB_EINVAL, "snb_pred_encode: B=%d beyond max_envs=%d", B, p->max_envs);
    SNB_REQUIRE(p->n_pushed > 0, SNB_EINVAL, "snb_pred_encode: no history (push or set the rings first)");
    SNB_REQUIRE(radius > 0.0 && dt > 0.0, SNB_EINVAL, "snb_pred_encode: radius and dt must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = run_encode(p, B, radius, dt, s);
    if (rc) return rc;
    const size_t BH = (size_t)B * p->H;
    if (ctx) SNB_CUDA_TRY(cudaMemcpyAsync(ctx, p->ctx, BH * 256 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (n_in) SNB_CUDA_TRY(cudaMemcpyAsync(n_in, p->prep.n_in, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (ped_ids) SNB_CUDA_TRY(cudaMemcpyAsync(ped_ids, p->prep.ped_ids, BH * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (in_cluster) SNB_CUDA_TRY(cudaMemcpyAsync(in_cluster, p->prep.in_cluster, BH, cudaMemcpyDeviceToDevice, s));
    return SNB_OK;
}

extern "C" int snb_pred_noise(float *out, int64_t n, uint64_t seed, uint64_t offset, void *stream)
{
    SNB_REQUIRE(out && n >= 0, SNB_EINVAL, "snb_pred_noise: bad argument");
    return snb_k_pred_noise(out, (size_t)n, seed, offset, (cudaStream_t)stream);
}

extern "C" int snb_pred_predict(SnbPredictor *p, int32_t B, const float *noise, uint64_t seed, int32_t n_steps, int32_t num_ret,
                                double radius, double dt, double *forecasts, double *logw, void *stream)
{
    SNB_REQUIRE(p && forecasts && logw, SNB_EINVAL, "snb_pred_predict: NULL argument");
    SNB_REQUIRE(B >= 0 && B <= p->max_envs, SNB_EINVAL, "snb_pred_predict: B=%d beyond max_envs=%d", B, p->max_envs);
    SNB_REQUIRE(num_ret >= 1 && num_ret <= p->S, SNB_EINVAL, "snb_pred_predict: num_ret=%d outside [1, S=%d]", num_ret, p->S);
    SNB_REQUIRE(p->n_pushed > 0, SNB_EINVAL, "snb_pred_predict: no history (push or set the rings first)");
    SNB_REQUIRE(radius > 0.0 && dt > 0.0, SNB_EINVAL, "snb_pred_predict: radius and dt must be positive");
    SNB_REQUIRE(n_steps >= 1 && n_steps <= 100 && 100 % (100 / n_steps) == 0, SNB_EINVAL,
                "snb_pred_predict: step_size=%d: int(100/step_size) must divide 100 (diffusion.py:507-537)", n_steps);
    if (B == 0) return SNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int H = p->H, S = p->S, T = p->T;
    int rc = run_encode(p, B, radius, dt, s);
    if (rc) return rc;
    SNB_CUDA_TRY(cudaMemcpyAsync(p->h_n_in, p->prep.n_in, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (!noise) {
        const size_t n = (size_t)B * S * H * T * 2;
        if ((rc = snb_k_pred_noise(p->noise, n, seed, p->calls * ((n + 3) / 4), s))) return rc;
        noise = p->noise;
    }
    ++p->calls;
    SNB_CUDA_TRY(cudaStreamSynchronize(s)); // the B cluster sizes decide the batch shapes of the denoiser
    // counting sort of the environments by cluster size (stable: ascending environment index inside a group)
    std::vector<int> cnt(H + 2, 0), off(H + 2, 0);
    for (int e = 0; e < B; ++e) {
        const int a = p->h_n_in[e];
        SNB_REQUIRE(a >= 1 && a <= H, SNB_ECUDA, "snb_pred_predict: cluster size %d of environment %d outside [1, %d]", a, e, H);
        ++cnt[a];
    }
    for (int a = 1; a <= H; ++a) off[a + 1] = off[a] + cnt[a];
    {
        std::vector<int> cur(off);
        for (int e = 0; e < B; ++e) p->h_order[cur[p->h_n_in[e]]++] = e;
    }
    SNB_CUDA_TRY(cudaMemcpyAsync(p->order, p->h_order, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    const bool kde = num_ret < S;
    for (int A = 1; A <= H; ++A) {
        const int c = cnt[A];
        if (!c) continue;
        const int32_t *ord = p->order + off[A];
        if ((rc = snb_k_pred_gather(ord, c, A, H, S, T, p->ctx, noise, p->ctx_b, p->xT_b, p->p0_b, p->prep.p0, s))) return rc;
        if ((rc = snb_jmid_denoise_agents(p->den, p->ctx_b, p->xT_b, p->vel_b, c, A, n_steps, s))) return rc;
        if ((rc = snb_jmid_integrate(p->vel_b, p->p0_b, p->pos_b, c, S, A, T, (float)dt, s))) return rc;
        if (kde && (rc = snb_k_pred_kde_topk(ord, c, A, S, T, num_ret, p->pos_b, p->sel, p->logw_env, p->kde_work, s))) return rc;
        if ((rc = snb_k_pred_scatter(ord, c, A, H, S, T, num_ret, p->pos_b, kde ? p->sel : nullptr, p->prep.ped_ids, forecasts, s))) return rc;
    }
    // np.log(np.ones(k) / k) (mid_sim_wrapper.py:491-493)
    return snb_k_pred_fill(B, H, T, num_ret, p->prep.in_cluster, p->prep.cv, p->prep.cur, std::log(1.0 / (double)num_ret),
                           kde ? p->logw_env : nullptr, forecasts, logw, s);
}

extern "C" int snb_pred_predict_host(SnbPredictor *p, const double *hist_host, const double *robot_hist_host, int32_t B,
                                     const float *noise_host, uint64_t seed, int32_t n_steps, int32_t num_ret, double radius, double dt,
                                     double *forecasts_host, double *logw_host)
{
    SNB_REQUIRE(p && hist_host && robot_hist_host && forecasts_host && logw_host, SNB_EINVAL, "snb_pred_predict_host: NULL argument");
    SNB_REQUIRE(B >= 1 && B <= p->max_envs, SNB_EINVAL, "snb_pred_predict_host: B=%d beyond max_envs=%d", B, p->max_envs);
    SNB_REQUIRE(num_ret >= 1 && num_ret <= p->S, SNB_EINVAL, "snb_pred_predict_host: num_ret=%d outside [1, S=%d]", num_ret, p->S);
    if (!p->own_stream) SNB_CUDA_TRY(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    cudaStream_t s = p->own_stream;
    const size_t fc_elems = (size_t)p->max_envs * p->H * p->S * (p->T + 1) * 2;
    if (!p->d_fc) {
        SNB_CUDA_TRY(cudaMalloc(&p->d_fc, fc_elems * sizeof(double)));
        SNB_CUDA_TRY(cudaMalloc(&p->d_lw, (size_t)p->max_envs * p->H * p->S * sizeof(double)));
    }
    int rc = snb_pred_set_history(p, hist_host, robot_hist_host, B, s);
    if (rc) return rc;
    const float *noise = nullptr;
    if (noise_host) {
        SNB_CUDA_TRY(cudaMemcpyAsync(p->noise, noise_host, (size_t)B * p->S * p->H * p->T * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
        noise = p->noise;
    }
    if ((rc = snb_pred_predict(p, B, noise, seed, n_steps, num_ret, radius, dt, p->d_fc, p->d_lw, s))) return rc;
    const size_t nf = (size_t)B * p->H * num_ret * (p->T + 1) * 2;
    SNB_CUDA_TRY(cudaMemcpyAsync(forecasts_host, p->d_fc, nf * sizeof(double), cudaMemcpyDeviceToHost, s));
    SNB_CUDA_TRY(cudaMemcpyAsync(logw_host, p->d_lw, (size_t)B * p->H * num_ret * sizeof(double), cudaMemcpyDeviceToHost, s));
    SNB_CUDA_TRY(cudaStreamSynchronize(s));
    return SNB_OK;
}

extern "C" int snb_pred_ingest(const double *forecasts, const double *logw, int32_t B, int32_t H, int32_t k, int32_t T, int32_t horiz,
                               double dt, int32_t joint, double *resh, double *weights, double *goals, double *vpref, void *stream)
{
    SNB_REQUIRE(forecasts && logw && resh && weights && goals && vpref, SNB_EINVAL, "snb_pred_ingest: NULL argument");
    SNB_REQUIRE(B >= 0 && H >= 1 && k >= 1 && T >= 2 && horiz >= 0 && dt > 0.0, SNB_EINVAL, "snb_pred_ingest: bad sizes");
    return snb_k_pred_ingest(forecasts, logw, B, H, k, T, horiz, dt, joint, resh, weights, goals, vpref, (cudaStream_t)stream);
}

extern "C" int snb_pred_mpc_pack(const double *robot, const double *humans, const double *goals, const double *weights, const double *resh,
                                 const double *stage_prefix, const double *static_obs, int32_t B, int32_t H, int32_t k, int32_t T,
                                 int32_t horiz, int32_t joint, int32_t n_prefix, int32_t n_static, double *mpc_state, double *human_theta,
                                 double *stage_params, void *stream)
{
    SNB_REQUIRE(robot && humans && goals && weights, SNB_EINVAL, "snb_pred_mpc_pack: NULL argument");
    SNB_REQUIRE(B >= 0 && H >= 1 && k >= 1 && horiz >= 1 && n_prefix >= 0 && n_static >= 0, SNB_EINVAL, "snb_pred_mpc_pack: bad sizes");
    SNB_REQUIRE(n_prefix == 0 || stage_prefix, SNB_EINVAL, "snb_pred_mpc_pack: n_prefix > 0 without stage_prefix");
    SNB_REQUIRE(n_static == 0 || static_obs, SNB_EINVAL, "snb_pred_mpc_pack: n_static > 0 without static_obs");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = SNB_OK;
    if (mpc_state) rc = snb_k_pred_mpc_state(robot, humans, goals, weights, B, H, k, joint, mpc_state, human_theta, s);
    if (!rc && stage_params) {
        SNB_REQUIRE(resh, SNB_EINVAL, "snb_pred_mpc_pack: stage_params needs the reshaped forecasts");
        // MID_samples[idx + 1] for idx = horiz - 1 must exist: forecasts_reshaped keeps min(T, horiz + 1) frames (:1651)
        SNB_REQUIRE(horiz + 1 <= T, SNB_EINVAL, "snb_pred_mpc_pack: horiz + 1 = %d frames needed, the forecasts hold T = %d", horiz + 1, T);
        rc = snb_k_pred_stage_params(resh, stage_prefix, static_obs, B, horiz, horiz + 1, H * k, n_prefix, n_static, stage_params, s);
    }
    return rc;
}

extern "C" int snb_env_log_push(const SnbCrowdState *st, double *log_dev, int32_t L, int32_t slot, void *stream)
{
    SNB_REQUIRE(st && log_dev, SNB_EINVAL, "snb_env_log_push: NULL argument");
    SNB_REQUIRE(L >= 1 && slot >= 0 && slot < L && st->E >= 1, SNB_EINVAL, "snb_env_log_push: bad ring slot / no robot");
    SNB_REQUIRE(st->E == 1, SNB_EINVAL, "snb_env_log_push: the simulator state has exactly one extra (the robot)");
    return snb_k_state_log_push(st->px, st->py, st->ex_px, st->ex_py, st->B, st->H, L, slot, log_dev, (cudaStream_t)stream);
}

extern "C" int snb_pred_bootstrap_history(SnbPredictor *p, const double *log_dev, int32_t L, int32_t newest, int32_t B, void *stream)
{
    SNB_REQUIRE(p && log_dev, SNB_EINVAL, "snb_pred_bootstrap_history: NULL argument");
    SNB_REQUIRE(B >= 0 && B <= p->max_envs, SNB_EINVAL, "snb_pred_bootstrap_history: B=%d beyond max_envs=%d", B, p->max_envs);
    SNB_REQUIRE(L >= SNB_PRED_TH + 1 && newest >= 0 && newest < L, SNB_EINVAL,
                "snb_pred_bootstrap_history: the log must hold past_num_frames + 1 = %d states (IndexError in the reference)", SNB_PRED_TH + 1);
    int rc = snb_k_pred_bootstrap(log_dev, B, p->H, L, newest, p->hist, p->robot_hist, (cudaStream_t)stream);
    if (!rc) p->n_pushed = SNB_PRED_TH;
    return rc;
}

extern "C" int snb_pred_kde_topk(const float *pos, int32_t B, int32_t S, int32_t A, int32_t T, int32_t k, int32_t *sel, double *logw,
                                 void *stream)
{
    SNB_REQUIRE(pos && sel && logw, SNB_EINVAL, "snb_pred_kde_topk: NULL argument");
    SNB_REQUIRE(B >= 0 && S >= 2 && A >= 1 && T >= 1 && k >= 1 && k <= S, SNB_EINVAL, "snb_pred_kde_topk: bad sizes");
    if (B == 0) return SNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    float *work = nullptr;
    SNB_CUDA_TRY(cudaMallocAsync(&work, snb_k_pred_kde_work_floats(B, A, S, T) * sizeof(float), s));
    int rc = snb_k_pred_kde_topk(nullptr, B, A, S, T, k, pos, sel, logw, work, s);
    SNB_CUDA_TRY(cudaFreeAsync(work, s));
    return rc;
}
