"""ctypes binding of libsnb.so (the C ABI declared in include/snb.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libsnb.so")


class SnbError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python safe-interactive-crowdnav_b200/build.py` "
        "(nvcc, sm_100a).  snb has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

OK, EINVAL, ECUDA, EUNSUPPORTED, ENOMEM, EOVERFLOW = 0, -1, -2, -3, -4, -5
POLICY_ORCA, POLICY_ORCA_PLUS, POLICY_SFM = 0, 1, 2
KIN_HOLONOMIC, KIN_UNICYCLE = 0, 1
F_REACHED, F_TIMEOUT, F_COLLISION, F_WALL, F_FROZEN, F_DANGER, F_DONE = 1, 2, 4, 8, 16, 32, 64
MAX_AGENTS_PER_ENV = 32
MAX_SEGMENTS = 64

_d, _i32, _vp = C.c_double, C.c_int32, C.c_void_p


class PolicyCfg(C.Structure):
    _fields_ = [("policy", _i32), ("max_neighbors", _i32), ("time_step", _d),
                ("neighbor_dist", _d), ("time_horizon", _d), ("time_horizon_obst", _d),
                ("policy_radius", _d), ("max_speed", _d), ("safety_space", _d),
                ("sfm_radius", _d), ("A", _d), ("B", _d), ("KI", _d), ("A_static", _d), ("B_static", _d),
                ("A_bottleneck", _d), ("B_bottleneck", _d), ("is_bottleneck", _i32), ("_pad", _i32)]


class DoorCfg(C.Structure):
    _fields_ = [("enabled", _i32), ("_pad", _i32), ("door_y_mid_min", _d), ("door_y_mid_max", _d), ("door_x_mid", _d),
                ("door_y_min", _d), ("door_y_max", _d), ("door_width", _d)]


class SceneCfg(C.Structure):
    _fields_ = [("rule", _i32), ("randomize_attributes", _i32), ("circle_radius", _d), ("rect_width", _d), ("rect_height", _d),
                ("human_radius", _d), ("human_v_pref", _d), ("robot_radius", _d), ("discomfort_dist", _d)]


SCENE_CIRCLE_CROSSING, SCENE_HALLWAY = 0, 1


class RewardCfg(C.Structure):
    _fields_ = [("success_reward", _d), ("timeout", _d), ("collision_penalty", _d), ("wall_collision_penalty", _d),
                ("freezing_penalty", _d), ("discomfort", _i32), ("has_progress", _i32), ("discomfort_dist", _d),
                ("discomfort_penalty_factor", _d), ("progress_factor", _d), ("time_limit", _d)]


class CrowdState(C.Structure):
    _fields_ = [("B", _i32), ("H", _i32), ("E", _i32), ("n_obs_extras", _i32)] + \
        [(n, _vp) for n in ("px", "py", "vx", "vy", "theta", "gx", "gy", "fgx", "fgy", "vpref", "radius", "human_time",
                            "ex_px", "ex_py", "ex_vx", "ex_vy", "ex_radius",
                            "rtheta", "rgx", "rgy", "global_time", "prev_dist")] + \
        [("robot_kinematics", _i32), ("_pad", _i32)]


class CslWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("layer_w", "layer_b", "hyper_bias_w", "hyper_gate_w", "hyper_gate_b")]


class EncLayerWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b", "lin2_w",
                                   "lin2_b", "norm1_w", "norm1_b", "norm2_w", "norm2_b")]


class JmidWeights(C.Structure):
    _fields_ = [("concat1", CslWeights), ("concat3", CslWeights), ("concat4", CslWeights), ("linear", CslWeights),
                ("layers", EncLayerWeights * 3), ("pos_emb", _vp), ("betas", _vp), ("alpha_bars", _vp)]


class LstmWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("w_ih", "w_hh", "b_ih", "b_hh")]


class EncoderWeights(C.Structure):
    _fields_ = [("node_history", LstmWeights), ("edge_ped", LstmWeights), ("edge_robot", LstmWeights),
                ("att_w1", _vp), ("att_w2", _vp), ("att_v", _vp)]


def _proto(name, restype, argtypes, required=True):
    try:
        f = getattr(lib, name)
    except AttributeError:
        if required:
            raise
        return None
    f.restype = restype
    f.argtypes = argtypes
    return f


_proto("snb_version", C.c_int, [])
_proto("snb_last_error", C.c_char_p, [])
_proto("snb_launch_count", C.c_uint64, [])
_proto("snb_obstacles_create", C.c_int, [C.POINTER(_vp), C.POINTER(_d), _i32])
_proto("snb_obstacles_destroy", C.c_int, [_vp])
_proto("snb_obstacles_num_vertices", _i32, [_vp])
_proto("snb_obstacles_get_vertex", C.c_int, [_vp, _i32, C.POINTER(C.c_float)])
_proto("snb_policy_step", C.c_int, [C.POINTER(PolicyCfg), C.POINTER(CrowdState), _vp, _vp, _vp, _vp, _vp, _vp])
_proto("snb_env_step", C.c_int, [C.POINTER(PolicyCfg), C.POINTER(DoorCfg), C.POINTER(RewardCfg), C.POINTER(CrowdState),
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp])
_proto("snb_scene_reset", C.c_int, [C.POINTER(SceneCfg), C.POINTER(DoorCfg), C.POINTER(CrowdState), _vp, _vp, _i32, _vp, _vp])
_proto("snb_env_whatif", C.c_int, [C.POINTER(PolicyCfg), C.POINTER(DoorCfg), C.POINTER(RewardCfg), C.POINTER(CrowdState),
                                    _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp])
_proto("snb_policy_predict_host", C.c_int, [C.POINTER(PolicyCfg), C.POINTER(_d), _i32, C.POINTER(_d), _i32, C.POINTER(_d),
                                             C.POINTER(_d), C.POINTER(_i32), C.POINTER(_i32)])
_proto("snb_robot_linear_action", C.c_int, [C.POINTER(CrowdState), _d, _vp, _vp])
_proto("snb_episode_metrics_update", C.c_int, [_vp, _vp, _vp, _vp, _d, _i32, _vp])
_proto("snb_jmid_create", C.c_int, [C.POINTER(_vp), C.POINTER(JmidWeights), _i32, _i32, _i32, _i32, _i32, _vp], required=False)
_proto("snb_jmid_destroy", C.c_int, [_vp], required=False)
_proto("snb_jmid_denoise", C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp], required=False)
_proto("snb_jmid_eps", C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp], required=False)
_proto("snb_jmid_integrate", C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, C.c_float, _vp], required=False)
_proto("snb_jmid_predict_host", C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.POINTER(C.c_float), _i32, _i32, C.c_float], required=False)
_proto("snb_jmid_gemm_bf16", C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp], required=False)
_proto("snb_jmid_attention", C.c_int, [_vp, _vp, _i32, _i32, _vp], required=False)
_proto("snb_jmid_flops_per_iter", _d, [_i32, _i32, _i32, _i32], required=False)
_proto("snb_jmid_denoise_agents", C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp], required=False)
_proto("snb_jmid_set_precision", C.c_int, [_vp, _i32, _vp], required=False)
_proto("snb_jmid_dims", C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)], required=False)
_u64 = C.c_uint64
_proto("snb_pred_create", C.c_int, [C.POINTER(_vp), C.POINTER(EncoderWeights), _vp, _i32, _i32, _vp], required=False)
_proto("snb_pred_destroy", C.c_int, [_vp], required=False)
_proto("snb_pred_push_history", C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp], required=False)
_proto("snb_pred_reset_history", C.c_int, [_vp], required=False)
_proto("snb_pred_set_history", C.c_int, [_vp, _vp, _vp, _i32, _vp], required=False)
_proto("snb_pred_encode", C.c_int, [_vp, _i32, _d, _d, _vp, _vp, _vp, _vp, _vp], required=False)
_proto("snb_pred_noise", C.c_int, [_vp, C.c_int64, _u64, _u64, _vp], required=False)
_proto("snb_pred_predict", C.c_int, [_vp, _i32, _vp, _u64, _i32, _i32, _d, _d, _vp, _vp, _vp], required=False)
_proto("snb_pred_predict_host", C.c_int, [_vp, C.POINTER(_d), C.POINTER(_d), _i32, C.POINTER(C.c_float), _u64, _i32, _i32, _d, _d,
                                           C.POINTER(_d), C.POINTER(_d)], required=False)
_proto("snb_pred_kde_topk", C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp], required=False)
_proto("snb_pred_ingest", C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _d, _i32, _vp, _vp, _vp, _vp, _vp], required=False)
_proto("snb_pred_set_position_std", C.c_int, [_vp, _d], required=False)
_proto("snb_pred_mpc_pack", C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
       required=False)
_proto("snb_env_step_logged", C.c_int, [C.POINTER(PolicyCfg), C.POINTER(DoorCfg), C.POINTER(RewardCfg), C.POINTER(CrowdState),
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp])
_proto("snb_env_log_push", C.c_int, [C.POINTER(CrowdState), _vp, _i32, _i32, _vp], required=False)
_proto("snb_pred_bootstrap_history", C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp], required=False)


def check(rc, what=""):
    if rc != 0:
        msg = lib.snb_last_error().decode("utf-8", "replace")
        raise SnbError(f"{what} failed with code {rc}: {msg}")


def launch_count():
    return int(lib.snb_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)
