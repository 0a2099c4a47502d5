/*
 * include/snb.h -- C ABI of libsnb.so, the B200-native replacement for the data-parallel hot path of
 * sepsamavi/safe-interactive-crowdnav (CrowdSimPlus per-agent ORCA/SFM step + JMID denoising loop).
 *
 * Boundary rules
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - pointers named *_dev are CUDA device pointers owned by the caller (e.g. torch.Tensor.data_ptr());
 *     pointers named *_host are ordinary host memory.  `stream` is a cudaStream_t passed as void* (NULL =
 *     the legacy default stream).  Device entry points are asynchronous with respect to `stream`.
 *   - every function returns 0 (SNB_OK) or a negative SNB_E* code and never throws; the message of the
 *     last failure on the calling thread is available from snb_last_error().
 *   - there is NO CPU fallback: if no CUDA device / kernel image is usable the call fails with SNB_ECUDA.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference root).
 */
#ifndef SNB_H
#define SNB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_VERSION 100 /* 0.1.0 */

enum {
    SNB_OK = 0,
    SNB_EINVAL = -1,       /* bad argument (NULL, size, alignment) */
    SNB_ECUDA = -2,        /* CUDA runtime / driver error, no device, launch failure */
    SNB_EUNSUPPORTED = -3, /* size beyond a compiled limit (see SNB_MAX_*) */
    SNB_ENOMEM = -4,
    SNB_EOVERFLOW = -5     /* a device-side capacity (ORCA lines, obstacle neighbours) was exceeded */
};

int snb_version(void);
const char *snb_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches counter) */
uint64_t snb_launch_count(void);

/* ======================================================================================================
 * Crowd step  (crowd_sim_plus/envs/crowd_sim_plus.py:1025-1257, envs/policy/{orca,orca_plus,social_force}.py)
 * ====================================================================================================== */

enum { SNB_POLICY_ORCA = 0, SNB_POLICY_ORCA_PLUS = 1, SNB_POLICY_SFM = 2 };
enum { SNB_KIN_HOLONOMIC = 0, SNB_KIN_UNICYCLE = 1 };

#define SNB_MAX_AGENTS_PER_ENV 32 /* humans + observed extras handled by one warp */
#define SNB_MAX_ORCA_LINES 32
#define SNB_MAX_SEGMENTS 64

/* Attributes of the reference policy objects: ORCA.__init__ (orca.py:55-67), ORCAPlus.configure
 * (orca_plus.py:15-27), SFM.configure (social_force.py:21-36), plus [env] time_step. */
typedef struct SnbPolicyCfg {
    int32_t policy;      /* SNB_POLICY_* */
    int32_t max_neighbors;
    double time_step;
    double neighbor_dist, time_horizon, time_horizon_obst;
    double policy_radius; /* PyRVOSimulator default radius (never used by an agent) */
    double max_speed;     /* max speed given to every OTHER agent of the throw-away simulator */
    double safety_space;
    double sfm_radius, A, B, KI, A_static, B_static, A_bottleneck, B_bottleneck;
    int32_t is_bottleneck;
    int32_t _pad;
} SnbPolicyCfg;

/* Human.get_g_xy door logic (envs/utils/human_plus.py:19-52) */
typedef struct SnbDoorCfg {
    int32_t enabled, _pad;
    double door_y_mid_min, door_y_mid_max, door_x_mid, door_y_min, door_y_max, door_width;
} SnbDoorCfg;

/* reward table of CrowdSimPlus.configure (crowd_sim_plus.py:87-128) */
typedef struct SnbRewardCfg {
    double success_reward, timeout, collision_penalty, wall_collision_penalty, freezing_penalty;
    int32_t discomfort, has_progress;
    double discomfort_dist, discomfort_penalty_factor, progress_factor, time_limit;
} SnbRewardCfg;

/* flag bits written per environment by snb_env_step (the non-zero info[...] entries of step()) */
enum {
    SNB_F_REACHED = 1, SNB_F_TIMEOUT = 2, SNB_F_COLLISION = 4, SNB_F_WALL = 8,
    SNB_F_FROZEN = 16, SNB_F_DANGER = 32, SNB_F_DONE = 64
};

/*
 * SoA state of B environments in HBM, all fp64 (the reference's Python floats; ORCA narrows to fp32 exactly
 * where Python-RVO2 does).  Humans: [B*H] arrays, element b*H+i.  "Extras" are the agents a human observes
 * after the other humans, in `ob` order (crowd_sim_plus.py:1047-1049): [B*E] arrays.  In the simulator E=1
 * and extra 0 IS the robot (its px/py/vx/vy are updated by snb_env_step); in the B=1 plugin call E = number
 * of observed agents.  Replaces the FullState / ObservableState / JointState objects
 * (envs/utils/state_plus.py:1-66).
 */
typedef struct SnbCrowdState {
    int32_t B, H, E;
    int32_t n_obs_extras;          /* how many extras the humans see (0 = robot invisible) */
    double *px, *py, *vx, *vy, *theta, *gx, *gy, *fgx, *fgy, *vpref, *radius, *human_time; /* [B*H] */
    double *ex_px, *ex_py, *ex_vx, *ex_vy, *ex_radius;                                     /* [B*E] */
    double *rtheta, *rgx, *rgy, *global_time, *prev_dist;                                  /* [B] (env step only) */
    int32_t robot_kinematics, _pad;
} SnbCrowdState;

/* Static line-segment obstacles shared by all environments (crowd_sim_plus.py:322-422).  Creation runs the
 * RVO2 obstacle pre-processing (addObstacle + processObstacles, orca_plus.py:50-53) ONCE on the host and
 * uploads the vertex list + BSP tree; the reference redoes it per human per step. */
typedef struct SnbObstacles SnbObstacles;
int snb_obstacles_create(SnbObstacles **out, const double *segs_host /* [n_seg*4] x1,y1,x2,y2 */, int32_t n_seg);
int snb_obstacles_destroy(SnbObstacles *obs);
int32_t snb_obstacles_num_vertices(const SnbObstacles *obs);
/* out7 = point.x point.y unitDir.x unitDir.y next prev isConvex (float), for parity tests of the BSP split */
int snb_obstacles_get_vertex(const SnbObstacles *obs, int32_t i, float *out7_host);

/*
 * Batched seeded scenario reset (CrowdSimPlus.reset -> generate_random_human_position, crowd_sim_plus.py:609-764, 425-451;
 * generate_circle_crossing_human :454-481; generate_hallway_human :522-605; Human.get_g_xy human_plus.py:19-52).
 * One thread per environment consumes the PCG64 stream of numpy's default_rng(seeds[b]) (SeedSequence hashing included) in the
 * reference's draw order and writes every human array of `state`, the robot at (0,-R) -> (0,R), rtheta = pi/2, clocks 0.
 * rule: SNB_SCENE_CIRCLE_CROSSING, or SNB_SCENE_HALLWAY for every rule generate_hallway_human serves (hallway,
 * hallway_static[_with_back], hallway_bottleneck, hallway_squeeze, rectangle, left_wall, no_walls; the segment list and the
 * door configuration carry the difference).  seeds_dev [B] uint64 = case offset + test case (:658-664).
 * segs_dev [n_seg*4] x1,y1,x2,y2 on the device.  n_draws_dev (optional int32 [B]) receives the number of 64-bit draws used,
 * or -1 where the rejection sampling gave up after 200 000 tries (an over-crowded scene; the reference would spin forever).
 */
#define SNB_SCENE_CIRCLE_CROSSING 0
#define SNB_SCENE_HALLWAY 1
typedef struct SnbSceneCfg {
    int32_t rule, randomize_attributes;
    double circle_radius, rect_width, rect_height, human_radius, human_v_pref, robot_radius, discomfort_dist;
} SnbSceneCfg;
int snb_scene_reset(const SnbSceneCfg *cfg, const SnbDoorCfg *door /* may be NULL */, const SnbCrowdState *state,
                    const uint64_t *seeds_dev, const double *segs_dev, int32_t n_seg, int32_t *n_draws_dev, void *stream);

/*
 * Human policy for every human of every environment, no clamp / integration:
 *   ORCA.predict (orca.py:82-133), ORCAPlus.predict (orca_plus.py:29-90), SFM.predict (social_force.py:38-94)
 * batched over B*H agents.  out_v_dev [B*H*2] (vx,vy).  Optional (ORCA only): nbr_dev [B*H*max_neighbors]
 * agent ids in RVO2 neighbour order (human j -> j, extra e -> H+e; -1 padded) and nbr_cnt_dev [B*H].
 * status_dev (optional, int32[1]) receives a non-zero SNB_EOVERFLOW marker if a device capacity was exceeded.
 */
int snb_policy_step(const SnbPolicyCfg *cfg, const SnbCrowdState *state, const SnbObstacles *obs /* may be NULL */,
                    double *out_v_dev, int32_t *nbr_dev, int32_t *nbr_cnt_dev, int32_t *status_dev, void *stream);

/*
 * One CrowdSimPlus.step(action, update=True) for B environments in ONE launch: human policies, static-obstacle
 * clamp (constrain_agent_action_exact, crowd_sim_plus.py:869-989), robot clamp + wall flag, robot-human
 * collision scan, frozen / goal / time-out, reward, state integration (Agent.step agent_plus.py:199-214,
 * Human.step human_plus.py:118-120), clocks and human arrival times.  State is updated in place.
 * robot_action_dev [B*2] = (vx,vy) or (v,r).  active_dev (optional uint8[B]): environments with 0 are skipped (state untouched)
 * and their reward / flags outputs are written as 0.
 * Outputs (each optional): reward_dev[B], dmin_dev[B], flags_dev[B] (SNB_F_*), nbr_dev / nbr_cnt_dev as above.
 */
int snb_env_step(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *reward_cfg,
                 const SnbCrowdState *state, const SnbObstacles *obs, const double *robot_action_dev,
                 const uint8_t *active_dev, double *reward_dev, double *dmin_dev, int32_t *flags_dev,
                 int32_t *nbr_dev, int32_t *nbr_cnt_dev, int32_t *status_dev, void *stream);

/*
 * CrowdSimPlus.step(action, update=False) -- the one-step look-ahead the RL observation builders call once per
 * discrete action (crowd_sim_plus.py:797-866 -> step(update=False), :1239-1255) -- for n_actions candidate robot
 * actions per environment in ONE launch.  The humans react to the current state only, so their policy (ORCA / SFM)
 * and clamp run once per environment; every candidate action then gets its own robot clamp, collision scan and
 * reward.  Nothing in `state` is modified (prev_dist included, crowd_sim_plus.py:1136-1137).
 *   robot_actions_dev [B, n_actions, 2];  outputs (each optional) reward_dev / dmin_dev / flags_dev [B, n_actions],
 *   next_humans_dev [B, H, 4] = Agent.get_next_observable_state (agent_plus.py:86-98) px,py,vx,vy of every human,
 *   next_robot_dev [B, n_actions, 2] = the robot's constrained next position.
 */
int snb_env_whatif(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *reward_cfg,
                   const SnbCrowdState *state, const SnbObstacles *obs, const double *robot_actions_dev,
                   int32_t n_actions, const uint8_t *active_dev, double *reward_dev, double *dmin_dev,
                   int32_t *flags_dev, double *next_humans_dev, double *next_robot_dev, int32_t *status_dev, void *stream);

/*
 * Host-buffer form of one policy call -- the literal replacement of what `policy.predict(state)` does behind
 * the rvo2 FFI (PyRVOSimulator(...) / addAgent / setAgentPrefVelocity / doStep / getAgentVelocity,
 * orca.py:95-129) or inside SFM.predict: copies the JointState to the device, runs the same kernel with
 * B=1,H=1,E=n_others, copies the action back, synchronises.
 *   self8 = px,py,vx,vy,radius,gx,gy,v_pref ; others5 = n x (px,py,vx,vy,radius) ; segs = m x (x1,y1,x2,y2)
 */
int snb_policy_predict_host(const SnbPolicyCfg *cfg, const double *self8_host, int32_t n_others,
                            const double *others5_host, int32_t n_seg, const double *segs_host,
                            double *out_v2_host, int32_t *nbr_ids_host /* [max_neighbors] ob indices, may be NULL */,
                            int32_t *n_nbr_host);

/*
 * Episode rollout glue (the test loop of simple_test.py:216-269 around CrowdSimPlus.step), device-resident:
 *   snb_robot_linear_action      the Linear robot policy (crowd_sim_plus/envs/policy/linear.py:16-23): action_dev [B,2] = v_pref * unit
 *                                vector from the robot (extra 0) to its goal.  Stand-in for the MPC solve, which is CPU code.
 *   snb_episode_metrics_update   per-environment episode counters from one step's flags / dmin (what simple_test.py accumulates from
 *                                `info`, :232-258, and pickles, :306-319): metrics_dev [B,9] fp64 = success, timeout, n_steps, nav_time,
 *                                n_collisions, n_wall_collisions, n_frozen, n_too_close, min_dist; live_dev [B] uint8 is cleared when an
 *                                environment reports SNB_F_DONE (finished environments stop counting).
 */
int snb_robot_linear_action(const SnbCrowdState *state, double v_pref, double *action_dev, void *stream);
int snb_episode_metrics_update(double *metrics_dev, uint8_t *live_dev, const int32_t *flags_dev, const double *dmin_dev,
                               double time_step, int32_t B, void *stream);

/* ======================================================================================================
 * JMID / iMID denoiser  (sicnav_diffusion/JMID/MID/models/diffusion.py:153-209, 478-541)
 * ====================================================================================================== */

/* fp32 device pointers into the reference state_dict; layouts as stored by torch
 * (nn.Linear weight = [out,in] row-major). */
typedef struct SnbCslWeights { /* ConcatSquashLinear, models/common.py:58-72 */
    const float *layer_w, *layer_b, *hyper_bias_w, *hyper_gate_w, *hyper_gate_b;
} SnbCslWeights;

typedef struct SnbEncLayerWeights { /* nn.TransformerEncoderLayer(512, 4, 1024), post-norm */
    const float *in_proj_w, *in_proj_b, *out_proj_w, *out_proj_b, *lin1_w, *lin1_b, *lin2_w, *lin2_b,
        *norm1_w, *norm1_b, *norm2_w, *norm2_b;
} SnbEncLayerWeights;

typedef struct SnbJmidWeights {
    SnbCslWeights concat1, concat3, concat4, linear;
    SnbEncLayerWeights layers[3];
    const float *pos_emb; /* [>=T,512] rows of net.pos_emb.pe */
    const float *betas, *alpha_bars; /* [101] var_sched buffers */
} SnbJmidWeights;

typedef struct SnbJmid SnbJmid;

/* Builds the device-resident model: converts the GEMM weights to bf16, allocates every activation buffer for
 * up to `max_envs` environments of A agents x S samples x T steps (no allocation happens afterwards).
 * joint=1: JointPredictionTransformerConcatLinear (one sequence of T*A*S tokens per env, quirk q1);
 * joint=0: TransformerConcatLinear (A*S sequences of T tokens).   Replaces MID._build_model (mid.py:1270-1297). */
int snb_jmid_create(SnbJmid **out, const SnbJmidWeights *w_dev, int32_t max_envs, int32_t A, int32_t S, int32_t T,
                    int32_t joint, void *stream);
int snb_jmid_destroy(SnbJmid *h);

/*
 * DiffusionTraj.sample_sicnav_inference (diffusion.py:478-541), DDIM, batched over B environments:
 *   ctx_dev  [B,A,256] fp32   context of each agent (Trajectron encoder output)
 *   x_T_dev  [B,S*A,T,2] fp32 initial noise, row r = s*A + a (injected; the reference draws torch.randn, quirk q2)
 *   out_vel_dev [B,S,A,T,2] fp32 velocities x_0
 * n_steps = the yaml `step_size` (stride = int(100/n_steps), t = 100, 100-stride, ..., stride).
 */
int snb_jmid_denoise(SnbJmid *h, const float *ctx_dev, const float *x_T_dev, float *out_vel_dev, int32_t B,
                     int32_t n_steps, void *stream);
/* agents per environment chosen per call: A <= the handle's A; ctx [B,A,256], x_T [B,S*A,T,2], out [B,S,A,T,2].
 * (the predictor's attention cluster changes size from step to step, mid_sim_wrapper.py:322-355) */
int snb_jmid_denoise_agents(SnbJmid *h, const float *ctx_dev, const float *x_T_dev, float *out_vel_dev, int32_t B, int32_t A,
                            int32_t n_steps, void *stream);
/* Arithmetic of the noise network.  SNB_PREC_BF16 (default): bf16 tensor-core operands, fp32 accumulation / softmax / LayerNorm
 * statistics / DDIM state.  SNB_PREC_FP32X: fp32-class -- fp32 activations, every nn.Linear as a split-bf16 (3 pieces, 6 partial
 * products) tcgen05 GEMM accumulated in fp32, attention / LayerNorm / ConcatSquash in fp32 on the CUDA cores -- the instrument that
 * separates rounding error from algorithmic error against the reference's fp32 torch path (models/diffusion.py:173-209); ~6x the
 * tensor FLOPs, small chunks.  Buffers for it are allocated on the first switch. */
enum { SNB_PREC_BF16 = 0, SNB_PREC_FP32X = 1 };
int snb_jmid_set_precision(SnbJmid *h, int32_t precision, void *stream);
/* sizes the handle was created with */
int snb_jmid_dims(const SnbJmid *h, int32_t *A, int32_t *S, int32_t *T, int32_t *joint);

/* one noise-network forward at diffusion step t (parity of JointPredictionTransformerConcatLinear.forward) */
int snb_jmid_eps(SnbJmid *h, const float *ctx_dev, const float *x_t_dev, float *eps_dev, int32_t B, int32_t t,
                 void *stream);
/* SingleIntegrator.integrate_samples (single_integrator.py:290-321): pos = cumsum_t(v)*dt + p0[a].
 * vel_dev [B,S,A,T,2], p0_dev [B,A,2] -> pos_dev [B,S,A,T,2] */
int snb_jmid_integrate(const float *vel_dev, const float *p0_dev, float *pos_dev, int32_t B, int32_t S, int32_t A,
                       int32_t T, float dt, void *stream);
/* host-buffer form: H2D(ctx, x_T) -> denoise -> integrate -> D2H(pos), synchronous (the predictor plugin path) */
int snb_jmid_predict_host(SnbJmid *h, const float *ctx_host, const float *x_T_host, const float *p0_host,
                          float *pos_host, int32_t B, int32_t n_steps, float dt);
/* Component-level entry points (parity tests and micro-benchmarks of the two tensor-core kernels):
 *   snb_jmid_gemm_bf16:  out[M,N] = A[M,K] * W[N,K]^T + bias[N]; A, W bf16 device buffers (row-major, K contiguous);
 *                        epi = 0: bf16 out, 1: ReLU then bf16 out, 2: fp32 out (the pre-LayerNorm form)
 *   snb_jmid_attention:  multi-head self-attention (4 heads x 128) over n_env independent unmasked sequences of
 *                        n_tok tokens; qkv [n_env, n_tok, 1536] bf16 (Q|K|V) -> out [n_env*n_tok, 512] bf16
 * (the torch call sites they replace: nn.Linear / nn.MultiheadAttention inside nn.TransformerEncoderLayer,
 *  models/diffusion.py:161-166) */
int snb_jmid_gemm_bf16(const void *A_dev, const void *W_dev, const float *bias_dev, void *out_dev, int32_t M, int32_t N,
                       int32_t K, int32_t epi, void *stream);
int snb_jmid_attention(const void *qkv_dev, void *out_dev, int32_t n_env, int32_t n_tok, void *stream);
/* algorithmic FLOPs of one denoise iteration for one environment */
double snb_jmid_flops_per_iter(int32_t A, int32_t S, int32_t T, int32_t joint);

/* ======================================================================================================
 * JMID predictor around the denoiser: what HumanTrajectoryForecasterSim.predict_ret_best does per call
 * (sicnav_diffusion/JMID/mid_sim_wrapper.py:172-509), batched over B environments of H humans.
 * ====================================================================================================== */

typedef struct SnbLstmWeights { /* torch nn.LSTM(input, 128), one layer: weight_ih_l0 [512,input], weight_hh_l0 [512,128], gates i,f,g,o */
    const float *w_ih, *w_hh, *b_ih, *b_hh;
} SnbLstmWeights;

typedef struct SnbEncoderWeights { /* checkpoint["encoder"], SURVEY Appendix B */
    SnbLstmWeights node_history; /* PEDESTRIAN/node_history_encoder              LSTM(6 -> 128)  */
    SnbLstmWeights edge_ped;     /* PEDESTRIAN->PEDESTRIAN/edge_encoder          LSTM(12 -> 128) */
    SnbLstmWeights edge_robot;   /* PEDESTRIAN->JRDB_ROBOT/edge_encoder          LSTM(12 -> 128) */
    const float *att_w1, *att_w2, *att_v; /* PEDESTRIAN/edge_influence_encoder  w1 [128,128], w2 [128,128], v [1,128] */
} SnbEncoderWeights;

typedef struct SnbPredictor SnbPredictor;

/* Device-resident predictor for up to max_envs environments of H humans (H <= 31, H <= the denoiser's A).
 * `denoiser` is borrowed (not destroyed with the predictor).  Replaces HumanTrajectoryForecasterSim.__init__ /
 * _init_MID (mid_sim_wrapper.py:207-241). */
int snb_pred_create(SnbPredictor **out, const SnbEncoderWeights *w_dev, SnbJmid *denoiser, int32_t max_envs, int32_t H,
                    void *stream);
int snb_pred_destroy(SnbPredictor *p);

/* update_state_hists (mid_sim_wrapper.py:185-204): appends one frame of positions to the 6-frame rings; the frames must be
 * dt apart (the reference's resampling, :283-310, is then the identity).  human_p{x,y}_dev [B,H], robot_p{x,y}_dev [B].
 * The first push after create / reset fills every frame of the ring. */
int snb_pred_push_history(SnbPredictor *p, const double *human_px_dev, const double *human_py_dev, const double *robot_px_dev,
                          const double *robot_py_dev, int32_t B, void *stream);
int snb_pred_reset_history(SnbPredictor *p);
/* overwrite the rings: hist_dev [B,H,6,2], robot_hist_dev [B,6,2] (oldest frame first) */
int snb_pred_set_history(SnbPredictor *p, const double *hist_dev, const double *robot_hist_dev, int32_t B, void *stream);

/* convert_to_mid_state_env + get_timesteps_data + Trajectron.get_latent (mid_sim_wrapper.py:313-437, preprocessing.py:428-694,
 * mgcvae.py:505-880) on the rings.  Outputs (each optional): ctx_dev [B,H,256] (slot a < n_in[b] = a-th in-cluster human in
 * ascending id), n_in_dev [B], ped_ids_dev [B,H] (slot -> human, -1 unused), in_cluster_dev [B,H]. */
int snb_pred_encode(SnbPredictor *p, int32_t B, double radius, double dt, float *ctx_dev, int32_t *n_in_dev,
                    int32_t *ped_ids_dev, uint8_t *in_cluster_dev, void *stream);

/* The reference standardises positions by the attention radius (std[0:2] = attention_radius, preprocessing.py:477-478, 540), and
 * that is what snb_pred_encode / snb_pred_predict do with their `radius` argument.  A benchmark that widens `radius` only to force
 * every human into the cluster (configs[3]: A = H) can pin the position scale the network was trained with (3.0) here;
 * pos_std = 0 restores the reference behaviour. */
int snb_pred_set_position_std(SnbPredictor *p, double pos_std);

/* standard-normal noise (Philox4x32-10 + Box-Muller), element i a function of (seed, offset, i) only */
int snb_pred_noise(float *out_dev, int64_t n, uint64_t seed, uint64_t offset, void *stream);

/*
 * predict_ret_best (mid_sim_wrapper.py:482-509) for B environments:
 *   noise_dev [B,S,H,T,2] fp32 or NULL (then drawn from (seed, call counter)); environment b with A_b in-cluster humans uses
 *   noise[b, s, a < A_b] as row s*A_b + a of the reference's x_T.
 *   num_ret <= S samples are returned; num_ret < S selects them with the KDE top-k (get_most_likely_samples, :14-169).
 *   forecasts_dev [B,H,num_ret,T+1,2] fp64 (frame 0 = current pose, constant-velocity rows outside the cluster),
 *   logw_dev [B,H,num_ret] fp64.
 * Environments are grouped by A_b and each group is denoised as one batch (one host synchronisation per call to read
 * the B cluster sizes).
 */
int snb_pred_predict(SnbPredictor *p, int32_t B, const float *noise_dev, uint64_t seed, int32_t n_steps, int32_t num_ret,
                     double radius, double dt, double *forecasts_dev, double *logw_dev, void *stream);
/* host-buffer form (the plugin call): histories in, forecasts out, synchronous */
int snb_pred_predict_host(SnbPredictor *p, const double *hist_host, const double *robot_hist_host, int32_t B,
                          const float *noise_host, uint64_t seed, int32_t n_steps, int32_t num_ret, double radius, double dt,
                          double *forecasts_host, double *logw_host);

/* Component-level entry of the KDE top-k (get_most_likely_samples, mid_sim_wrapper.py:14-169; always the joint branch, quirk q4):
 * pos_dev [B,S,A,T,2] fp32 -> sel_dev [B,k] sample indices in ascending total log-likelihood (ties: ascending index),
 * logw_dev [B,k] fp64 = log-softmax of the kept totals. */
int snb_pred_kde_topk(const float *pos_dev, int32_t B, int32_t S, int32_t A, int32_t T, int32_t k, int32_t *sel_dev,
                      double *logw_dev, void *stream);

/* MPC ingest (SICNavAcados.predict, sicnav_diffusion/policy/sicnav_acados.py:1645-1667):
 *   resh_dev [B, min(T, horiz+1), H*k, 2]   'h s t d -> t (h s) d' of the forecasts without the current pose
 *   weights_dev [B,k] (joint: logw[b,0,:]) or [B,H,k];  goals_dev [B,H,2] = mean over samples of the first predicted point;
 *   vpref_dev [B,H] = max over samples and steps of |dp| / dt */
int snb_pred_ingest(const double *forecasts_dev, const double *logw_dev, int32_t B, int32_t H, int32_t k, int32_t T,
                    int32_t horiz, double dt, int32_t joint, double *resh_dev, double *weights_dev, double *goals_dev,
                    double *vpref_dev, void *stream);

/* What SICNavAcados.predict hands the solver after the ingest (sicnav_diffusion/policy/sicnav_acados.py), batched over B envs:
 *   mpc_state_dev [B, 10 + nX_hums] = convert_to_mpc_state_vector (:222-289) of the joint state built at :1655-1681:
 *       px, py, sin(theta), cos(theta), v, omega, v_dot, omega_dot, gx, gy | per human px, py, vx, vy, gx, gy (+ k log-weights when
 *       joint = 0, iMID) | k log-weights (joint = 1, JMID);  nX_hums = 6 H + k (joint) or (6 + k) H.
 *       robot_dev [B,9] = px, py, theta, lvel, omega, v_dot, omega_dot, gx, gy;  humans_dev [B,H,4] = px, py, vx, vy;
 *       goals_dev [B,H,2], weights_dev as written by snb_pred_ingest.  human_theta_dev [B,H] (optional) = atan2(vy, vx), 0 at rest (:1678).
 *   stage_params_dev [B, horiz+1, n_prefix + 4 H k + n_static] = the per-stage vector of :1389-1413:
 *       prefix (x_ref, u_ref, Q, R, Q_T: MPC-side, stage_prefix_dev [B, horiz+1, n_prefix]) | X_t[:,0] | X_t[:,1] | X_t+1[:,0] | X_t+1[:,1]
 *       (X = resh_dev [B, horiz+1, H k, 2] of snb_pred_ingest; stage horiz reuses t = horiz - 1, :1403-1405) | static_obs_dev [n_static].
 * mpc_state_dev / stage_params_dev may each be NULL. */
int snb_pred_mpc_pack(const double *robot_dev, const double *humans_dev, const double *goals_dev, const double *weights_dev,
                      const double *resh_dev, const double *stage_prefix_dev, const double *static_obs_dev, int32_t B, int32_t H,
                      int32_t k, int32_t T, int32_t horiz, int32_t joint, int32_t n_prefix, int32_t n_static, double *mpc_state_dev,
                      double *human_theta_dev, double *stage_params_dev, void *stream);

/* History bootstrap of reset_scenario_values (sicnav_acados.py:1163-1182): the forecaster is re-created and fed
 * env.states[-past_num_frames-1 : -1] (the newest logged state is skipped; stamps are dt apart).
 *   snb_env_log_push: CrowdSimPlus.step's `self.states.append` (crowd_sim_plus.py:1175-1181) -- writes the positions of `state`
 *       (humans + robot, BEFORE the step) into slot `slot` of log_dev [B, L, H+1, 2] (entry H = robot);
 *   snb_pred_bootstrap_history: fills the predictor's rings from the L-deep log whose newest slot is `newest`. */
int snb_env_log_push(const SnbCrowdState *state, double *log_dev, int32_t L, int32_t slot, void *stream);
/* snb_env_step with the log push folded into the same launch: ring slot `slot` of log_dev [B, L, H + 1, 2] receives the positions
 * BEFORE the step (what `self.states.append` holds for this step, crowd_sim_plus.py:1175-1181), of every environment, frozen or not. */
int snb_env_step_logged(const SnbPolicyCfg *cfg, const SnbDoorCfg *door, const SnbRewardCfg *reward_cfg,
                        const SnbCrowdState *state, const SnbObstacles *obs, const double *robot_action_dev,
                        const uint8_t *active_dev, double *reward_dev, double *dmin_dev, int32_t *flags_dev,
                        int32_t *nbr_dev, int32_t *nbr_cnt_dev, int32_t *status_dev, double *log_dev, int32_t L, int32_t slot,
                        void *stream);
int snb_pred_bootstrap_history(SnbPredictor *p, const double *log_dev, int32_t L, int32_t newest, int32_t B, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB_H */
