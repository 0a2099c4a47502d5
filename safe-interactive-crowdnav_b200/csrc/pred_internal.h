// csrc/pred_internal.h -- internal interfaces of the JMID predictor front / back end (pred_prep.cu, pred_encode.cu,
// pred_post.cu, pred_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "snb_common.h"

#define SNB_PRED_TH 6     // history frames kept per track (mid_sim_wrapper.py:199-202)
#define SNB_PRED_HID 128  // LSTM hidden size of every encoder

// Per-environment slot arrays: slot a < n_in[env] holds the a-th in-cluster human in ascending id.
struct PredPrepOut {
    int32_t *n_in;       // [B]
    uint8_t *in_cluster; // [B,H]
    int32_t *ped_ids;    // [B,H]   slot -> human index (-1 for unused slots)
    float *x_st;         // [B,H,6,6]  standardised own history
    float *nb_ped;       // [B,H,6,6]  sum of standardised PEDESTRIAN neighbour histories
    float *nb_rob;       // [B,H,6,6]  same for the JRDB_ROBOT edge type
    float *edge_mask;    // [B,H]      clamp(sum of edge scalings, <= 1)
    float *p0;           // [B,H,2]    current position (fp32, the integrator's initial condition)
    double *cv;          // [B,H,horizon,2] constant-velocity forecast of every human
    double *cur;         // [B,H,2]    current position (fp64)
};
int snb_k_pred_prep(const double *hist, const double *robot_hist, int B, int H, double radius, double pos_std, double dt, int horizon,
                    const PredPrepOut *out, cudaStream_t s);
int snb_k_pred_mpc_state(const double *robot, const double *humans, const double *goals, const double *weights, int B, int H, int k,
                         int joint, double *out, double *theta, cudaStream_t s);
int snb_k_pred_stage_params(const double *resh, const double *prefix, const double *stat, int B, int horiz, int Tp, int HK, int n_prefix,
                            int n_stat, double *out, cudaStream_t s);
int snb_k_state_log_push(const double *hpx, const double *hpy, const double *rpx, const double *rpy, int B, int H, int L, int slot,
                         double *log, cudaStream_t s);
int snb_k_pred_bootstrap(const double *log, int B, int H, int L, int newest, double *hist, double *robot_hist, cudaStream_t s);
int snb_k_pred_push(double *hist, double *robot_hist, const double *hpx, const double *hpy, const double *rpx, const double *rpy,
                    int B, int H, int first, cudaStream_t s);

// Encoder weights on the device, transposed for coalesced reads: w_ihT [din][512], w_hhT [128][512], bias [512] (= b_ih + b_hh)
struct PredLstmDev { const float *w_ihT, *w_hhT, *bias; int din; };
struct PredEncDev {
    PredLstmDev lstm[3];            // 0: node history (6), 1: PEDESTRIAN->PEDESTRIAN edge (12), 2: PEDESTRIAN->JRDB_ROBOT edge (12)
    const float *w1T, *w2T, *v;     // additive attention: w1T / w2T [128 in][128 out], v [128]
};
// ctx [rows, 256] = [attention-combined edge influence (128) | node history encoding (128)]
int snb_k_pred_encode(const PredEncDev *w, const float *x_st, const float *nb_ped, const float *nb_rob, const float *edge_mask,
                      float *ctx, int rows, cudaStream_t s);
int snb_k_transpose_f32(const float *src, float *dst, int rows, int cols, cudaStream_t s); // dst[c][r] = src[r][c]
int snb_k_add_f32(const float *a, const float *b, float *dst, int n, cudaStream_t s);

// ---- back end (pred_post.cu) ----
// standard-normal noise, Philox4x32-10 + Box-Muller, element i depends on (seed, i) only
int snb_k_pred_noise(float *out, size_t n, uint64_t seed, uint64_t offset, cudaStream_t s);
// bucket gather: envs order[off .. off+cnt) all have A agents;  ctx_b [cnt,A,256], xT_b [cnt,S*A,T,2] (row s*A+a)
// from ctx [B,H,256] and noise [B,S,H,T,2]
int snb_k_pred_gather(const int32_t *order, int cnt, int A, int H, int S, int T, const float *ctx, const float *noise, float *ctx_b,
                      float *xT_b, float *p0_b, const float *p0, cudaStream_t s);
// integrate the bucket's velocities (SingleIntegrator, fp32) -> pos_b [cnt,S,A,T,2]
// and, when sel == NULL, scatter sample s -> forecasts[env][ped][s][1+t] (k == S); with sel [cnt,k] scatter the selected samples.
int snb_k_pred_scatter(const int32_t *order, int cnt, int A, int H, int S, int T, int k, const float *pos_b, const int32_t *sel,
                       const int32_t *ped_ids, double *forecasts, cudaStream_t s);
// current pose at t = 0 of every (human, sample); constant-velocity rows for the humans outside the cluster;
// logw rows: uniform log(1/S) when logw_env == NULL else the environment's k cluster weights for every human
int snb_k_pred_fill(int B, int H, int T, int k, const uint8_t *in_cluster, const double *cv, const double *cur, double uniform_logw,
                    const double *logw_env, double *forecasts, double *logw, cudaStream_t s);
// KDE top-k (get_most_likely_samples, mid_sim_wrapper.py:14-169) for one bucket: pos_b [cnt,S,A,T,2] -> sel [cnt,k] sample
// indices in ascending total log-likelihood, logw_env[order[i]][k]
int snb_k_pred_kde_topk(const int32_t *order, int cnt, int A, int S, int T, int k, const float *pos_b, int32_t *sel, double *logw_env,
                        float *work, cudaStream_t s);
size_t snb_k_pred_kde_work_floats(int cnt, int A, int S, int T);
// MPC ingest (sicnav_acados.py:1645-1667)
int snb_k_pred_ingest(const double *forecasts, const double *logw, int B, int H, int k, int T, int horiz, double dt, int joint,
                      double *resh, double *weights, double *goals, double *vpref, cudaStream_t s);
