"""oracle/oracle_lib.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/liboracle.so (rvo2_oracle.c + crowd_oracle.c).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under safe-interactive-crowdnav_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("rvo2_oracle.c", "crowd_oracle.c", "rvo2_oracle.h", "crowd_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class PolicyCfg(C.Structure):
    _fields_ = [("policy", C.c_int), ("time_step", C.c_double),
                ("neighbor_dist", C.c_double), ("max_neighbors", C.c_int),
                ("time_horizon", C.c_double), ("time_horizon_obst", C.c_double),
                ("policy_radius", C.c_double), ("max_speed", C.c_double), ("safety_space", C.c_double),
                ("sfm_radius", C.c_double), ("A", C.c_double), ("B", C.c_double), ("KI", C.c_double),
                ("A_static", C.c_double), ("B_static", C.c_double),
                ("A_bottleneck", C.c_double), ("B_bottleneck", C.c_double), ("is_bottleneck", C.c_int)]


class DoorCfg(C.Structure):
    _fields_ = [("enabled", C.c_int), ("door_y_mid_min", C.c_double), ("door_y_mid_max", C.c_double),
                ("door_x_mid", C.c_double), ("door_y_min", C.c_double), ("door_y_max", C.c_double),
                ("door_width", C.c_double)]


class RewardCfg(C.Structure):
    _fields_ = [("success_reward", C.c_double), ("timeout", C.c_double), ("collision_penalty", C.c_double),
                ("wall_collision_penalty", C.c_double), ("freezing_penalty", C.c_double),
                ("discomfort", C.c_int), ("discomfort_dist", C.c_double), ("discomfort_penalty_factor", C.c_double),
                ("has_progress", C.c_int), ("progress_factor", C.c_double), ("time_limit", C.c_double)]


_DP = C.POINTER(C.c_double)


class EnvState(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int)] + \
        [(n, _DP) for n in ("px", "py", "vx", "vy", "theta", "gx", "gy", "fgx", "fgy", "vpref", "radius", "human_time",
                            "rpx", "rpy", "rvx", "rvy", "rtheta", "rgx", "rgy")] + \
        [("rradius", C.c_double), ("rvpref", C.c_double), ("robot_kinematics", C.c_int), ("robot_visible", C.c_int),
         ("global_time", _DP), ("prev_dist", _DP), ("n_seg", C.c_int), ("segs", _DP)]


PROBE_MAX_LINES = 48


class OrcProbe(C.Structure):
    _fields_ = [("n_lines", C.c_int), ("n_obst_lines", C.c_int), ("n_agent_nbr", C.c_int), ("nbr_ids", C.c_int * 16),
                ("lines", C.c_float * (4 * PROBE_MAX_LINES)), ("pref", C.c_float * 2), ("max_speed", C.c_float)]


POLICY_ORCA, POLICY_ORCA_PLUS, POLICY_SFM = 0, 1, 2
KIN_HOLONOMIC, KIN_UNICYCLE = 0, 1
F_REACHED, F_TIMEOUT, F_COLLISION, F_WALL, F_FROZEN, F_DANGER, F_DONE = 1, 2, 4, 8, 16, 32, 64

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        L.rvo_create.restype = C.c_void_p
        L.rvo_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        L.rvo_destroy.argtypes = [C.c_void_p]
        L.rvo_add_agent.restype = C.c_int
        L.rvo_add_agent.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        L.rvo_add_obstacle.restype = C.c_int
        L.rvo_add_obstacle.argtypes = [C.c_void_p, fp, C.c_int]
        L.rvo_process_obstacles.argtypes = [C.c_void_p]
        for n in ("rvo_set_agent_position", "rvo_set_agent_velocity", "rvo_set_agent_pref_velocity"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        for n in ("rvo_get_agent_position", "rvo_get_agent_velocity", "rvo_get_agent_pref_velocity"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, fp]
        L.rvo_get_agent_max_speed.restype = C.c_float
        L.rvo_get_agent_max_speed.argtypes = [C.c_void_p, C.c_int]
        L.rvo_get_num_agents.argtypes = [C.c_void_p]
        L.rvo_get_num_obstacle_vertices.argtypes = [C.c_void_p]
        L.rvo_get_global_time.restype = C.c_float
        L.rvo_get_global_time.argtypes = [C.c_void_p]
        L.rvo_do_step.argtypes = [C.c_void_p]
        for n in ("rvo_get_agent_num_agent_neighbors", "rvo_get_agent_num_obstacle_neighbors", "rvo_get_agent_num_orca_lines"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        for n in ("rvo_get_agent_agent_neighbor", "rvo_get_agent_obstacle_neighbor"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.rvo_get_agent_orca_line.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        L.rvo_get_obstacle_vertex.argtypes = [C.c_void_p, C.c_int, fp]

        L.orc_orca_predict.argtypes = [C.POINTER(PolicyCfg), _DP, C.c_int, _DP, C.c_int, _DP, _DP, ip, ip, ip, ip]
        L.orc_sfm_predict.argtypes = [C.POINTER(PolicyCfg), _DP, C.c_int, _DP, C.c_int, _DP, _DP]
        L.orc_closest_point_on_segment.argtypes = [C.c_double] * 6 + [_DP]
        L.orc_point_to_segment_dist.restype = C.c_double
        L.orc_point_to_segment_dist.argtypes = [C.c_double] * 6
        L.orc_closest_distance_between_line_segments.argtypes = [_DP] * 5
        L.orc_constrain_action.argtypes = [_DP, C.c_double, C.c_double, C.c_int, _DP, C.c_int, _DP, _DP]
        L.orc_env_step.argtypes = [C.POINTER(PolicyCfg), C.POINTER(DoorCfg), C.POINTER(RewardCfg), C.POINTER(EnvState),
                                   _DP, C.POINTER(C.c_ubyte), _DP, _DP, ip, ip, ip, C.c_int]
        L.orc_policy_batch.argtypes = [C.POINTER(PolicyCfg), C.POINTER(EnvState), _DP, ip, ip, C.c_int]
        L.orc_orca_probe.argtypes = [C.POINTER(PolicyCfg), C.c_int, _DP, C.c_int, _DP, ip, C.c_int, _DP, _DP, C.POINTER(OrcProbe)]
        L.rvo_branch_counters.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
        _lib = L
    return _lib


def dptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_DP)


def iptr(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


def default_policy_cfg(policy="orca", time_step=0.25, safety_space=0.0, **kw):
    """Constants of crowd_sim_plus/envs/policy/orca.py:55-67 and the shipped SFM section
    (sicnav_diffusion/configs/env.config:30-49)."""
    pol = {"orca": POLICY_ORCA, "orca_plus": POLICY_ORCA_PLUS, "sfm": POLICY_SFM}[policy]
    cfg = PolicyCfg(policy=pol, time_step=time_step, neighbor_dist=10.0, max_neighbors=10, time_horizon=2.0,
                    time_horizon_obst=0.5, policy_radius=0.3, max_speed=1.0, safety_space=safety_space,
                    sfm_radius=0.2, A=3.0, B=0.18, KI=1.0, A_static=2.0, B_static=0.025,
                    A_bottleneck=6.0, B_bottleneck=0.12, is_bottleneck=0)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def default_reward_cfg(time_limit=30.0, **kw):
    """sicnav_diffusion/configs/env.config:65-70 + the non-SB3 defaults of crowd_sim_plus.py:117-128."""
    cfg = RewardCfg(success_reward=1.0, timeout=-1.0, collision_penalty=-0.25, wall_collision_penalty=-1.0,
                    freezing_penalty=-0.125, discomfort=1, discomfort_dist=0.2, discomfort_penalty_factor=0.5,
                    has_progress=0, progress_factor=0.0, time_limit=time_limit)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


class EnvArrays:
    """fp64 SoA state of B environments x H humans (+ robot), the layout orc_env_step works on."""
    HUMAN = ("px", "py", "vx", "vy", "theta", "gx", "gy", "fgx", "fgy", "vpref", "radius", "human_time")
    ROBOT = ("rpx", "rpy", "rvx", "rvy", "rtheta", "rgx", "rgy")

    def __init__(self, B, H, segs=None, rradius=0.25, rvpref=1.0, robot_kinematics=KIN_HOLONOMIC, robot_visible=True):
        self.B, self.H = B, H
        for n in self.HUMAN:
            setattr(self, n, np.zeros(B * H, np.float64))
        for n in self.ROBOT:
            setattr(self, n, np.zeros(B, np.float64))
        self.global_time = np.zeros(B, np.float64)
        self.prev_dist = np.zeros(B, np.float64)
        self.segs = np.ascontiguousarray(np.zeros((0, 4)) if segs is None else np.asarray(segs, np.float64).reshape(-1, 4))
        self.rradius, self.rvpref = rradius, rvpref
        self.robot_kinematics, self.robot_visible = robot_kinematics, robot_visible

    def copy(self):
        o = EnvArrays(self.B, self.H, self.segs.copy(), self.rradius, self.rvpref, self.robot_kinematics, self.robot_visible)
        for n in self.HUMAN + self.ROBOT + ("global_time", "prev_dist"):
            getattr(o, n)[:] = getattr(self, n)
        return o

    def cstruct(self):
        s = EnvState(B=self.B, H=self.H, rradius=self.rradius, rvpref=self.rvpref,
                     robot_kinematics=self.robot_kinematics, robot_visible=int(self.robot_visible),
                     n_seg=len(self.segs))
        for n in self.HUMAN + self.ROBOT + ("global_time", "prev_dist"):
            setattr(s, n, dptr(getattr(self, n)))
        s.segs = dptr(self.segs) if len(self.segs) else C.cast(None, _DP)
        return s


def policy_batch(pcfg, env, n_threads=1, want_nbr=True):
    """Human policy only (no clamp/integrate): returns v[B,H,2] (+ nbr[B,H,MN], cnt[B,H])."""
    L = lib()
    B, H, MN = env.B, env.H, pcfg.max_neighbors
    out = np.zeros(B * H * 2, np.float64)
    nbr = np.full(B * H * MN, -1, np.int32)
    cnt = np.zeros(B * H, np.int32)
    st = env.cstruct()
    L.orc_policy_batch(C.byref(pcfg), C.byref(st), dptr(out), iptr(nbr) if want_nbr else None,
                       iptr(cnt) if want_nbr else None, n_threads)
    return out.reshape(B, H, 2), nbr.reshape(B, H, MN), cnt.reshape(B, H)


def env_step(pcfg, door, rcfg, env, robot_action, active=None, n_threads=1, want_nbr=False):
    """In-place CrowdSimPlus.step for every env; returns reward[B], dmin[B], flags[B] (+nbr, cnt)."""
    L = lib()
    B, H, MN = env.B, env.H, pcfg.max_neighbors
    ra = np.ascontiguousarray(np.asarray(robot_action, np.float64).reshape(B, 2))
    reward = np.zeros(B, np.float64)
    dmin = np.zeros(B, np.float64)
    flags = np.zeros(B, np.int32)
    nbr = np.full(B * H * MN, -1, np.int32) if want_nbr else None
    cnt = np.zeros(B * H, np.int32) if want_nbr else None
    st = env.cstruct()
    act = None if active is None else np.ascontiguousarray(active, np.uint8)
    L.orc_env_step(C.byref(pcfg), C.byref(door) if door is not None else None, C.byref(rcfg), C.byref(st), dptr(ra),
                   act.ctypes.data_as(C.POINTER(C.c_ubyte)) if act is not None else None,
                   dptr(reward), dptr(dmin), iptr(flags), iptr(nbr) if want_nbr else None,
                   iptr(cnt) if want_nbr else None, n_threads)
    if want_nbr:
        return reward, dmin, flags, nbr.reshape(B, H, MN), cnt.reshape(B, H)
    return reward, dmin, flags


def orca_probe(pcfg, self8, others, n_others, segs=None):
    """n independent ORCA(.Plus).predict calls: self8 [n,8], others [n,E,5], n_others [n] -> dict(v [n,2] float64 (float32 values),
    lines [n,PROBE_MAX_LINES,4] float32 (point, direction; obstacle lines first), n_lines, n_obst_lines, nbr [n,16], n_nbr,
    pref [n,2], max_speed [n])."""
    L = lib()
    self8 = np.ascontiguousarray(self8, np.float64); others = np.ascontiguousarray(others, np.float64)
    n, E = others.shape[0], others.shape[1]
    n_others = np.ascontiguousarray(n_others, np.int32)
    segs = np.zeros((0, 4)) if segs is None else np.ascontiguousarray(np.asarray(segs, np.float64).reshape(-1, 4))
    out_v = np.zeros((n, 2), np.float64)
    probes = (OrcProbe * n)()
    L.orc_orca_probe(C.byref(pcfg), n, dptr(self8.reshape(-1)), E, dptr(others.reshape(-1)), iptr(n_others), len(segs),
                     dptr(segs.reshape(-1)) if len(segs) else None, dptr(out_v.reshape(-1)), probes)
    raw = np.frombuffer(probes, dtype=np.dtype([("n_lines", "i4"), ("n_obst_lines", "i4"), ("n_agent_nbr", "i4"), ("nbr_ids", "i4", 16),
                                                ("lines", "f4", (PROBE_MAX_LINES, 4)), ("pref", "f4", 2), ("max_speed", "f4")]))
    assert raw.shape[0] == n
    return dict(v=out_v, lines=raw["lines"].copy(), n_lines=raw["n_lines"].copy(), n_obst_lines=raw["n_obst_lines"].copy(),
                nbr=raw["nbr_ids"].copy(), n_nbr=raw["n_agent_nbr"].copy(), pref=raw["pref"].astype(np.float64),
                max_speed=raw["max_speed"].astype(np.float64))


def branch_counters(reset=False):
    out = (C.c_ulonglong * 48)()
    lib().rvo_branch_counters(out, int(reset))
    return np.array(out[:], dtype=np.int64)
