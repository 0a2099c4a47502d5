// csrc/rollout_kernels.cu -- the per-step glue of an episode rollout, kept on the device so that a whole step is libsnb launches only:
//   * the stand-in robot policy Linear (crowd_sim_plus/envs/policy/linear.py:16-23; the MPC solve is CPU code outside the path),
//   * the per-environment episode counters simple_test.py accumulates from `info` (simple_test.py:216-269, 306-319).
#include "snb_common.h"

namespace {

__global__ void robot_linear_kernel(const double *__restrict__ rpx, const double *__restrict__ rpy, const double *__restrict__ rgx,
                                    const double *__restrict__ rgy, int B, int E, double v_pref, double *__restrict__ action)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double theta = atan2(rgy[b] - rpy[(size_t)b * E], rgx[b] - rpx[(size_t)b * E]);
    action[2 * b] = cos(theta) * v_pref;
    action[2 * b + 1] = sin(theta) * v_pref;
}

// columns: success, timeout, n_steps, nav_time, n_collisions, n_wall_collisions, n_frozen, n_too_close, min_dist
__global__ void episode_metrics_kernel(double *__restrict__ m, uint8_t *__restrict__ live, const int32_t *__restrict__ flags,
                                       const double *__restrict__ dmin, double dt, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !live[b]) return;
    const int f = flags[b];
    double *r = m + (size_t)b * 9;
    r[2] += 1.0;
    r[3] += dt;
    if (f & SNB_F_COLLISION) r[4] += 1.0;
    if (f & SNB_F_WALL) r[5] += 1.0;
    if (f & SNB_F_FROZEN) r[6] += 1.0;
    if (f & SNB_F_DANGER) r[7] += 1.0;
    const double d = dmin[b];
    if (d < r[8]) r[8] = d;
    if (f & SNB_F_REACHED) r[0] = 1.0;
    if (f & SNB_F_TIMEOUT) r[1] = 1.0;
    if (f & SNB_F_DONE) live[b] = 0;
}

} // namespace

extern "C" int snb_robot_linear_action(const SnbCrowdState *st, double v_pref, double *action_dev, void *stream)
{
    SNB_REQUIRE(st && action_dev, SNB_EINVAL, "snb_robot_linear_action: NULL argument");
    SNB_REQUIRE(st->E >= 1, SNB_EINVAL, "snb_robot_linear_action: the state has no robot (E = 0)");
    if (st->B == 0) return SNB_OK;
    robot_linear_kernel<<<(st->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(st->ex_px, st->ex_py, st->rgx, st->rgy, st->B, st->E, v_pref,
                                                                              action_dev);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}

extern "C" int snb_episode_metrics_update(double *metrics_dev, uint8_t *live_dev, const int32_t *flags_dev, const double *dmin_dev,
                                          double time_step, int32_t B, void *stream)
{
    SNB_REQUIRE(metrics_dev && live_dev && flags_dev && dmin_dev, SNB_EINVAL, "snb_episode_metrics_update: NULL argument");
    if (B <= 0) return SNB_OK;
    episode_metrics_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(metrics_dev, live_dev, flags_dev, dmin_dev, time_step, B);
    snb_count_launch();
    SNB_CUDA_TRY(cudaGetLastError());
    return SNB_OK;
}
