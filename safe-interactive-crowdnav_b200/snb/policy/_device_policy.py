"""Shared machinery of the CUDA-backed human policies: packs a JointState for the B=1 plugin call and builds the
SnbPolicyCfg the kernels read."""
import ctypes as C

import numpy as np

from .. import _capi
from ..utils.action import ActionXY


def policy_cfg(policy_obj, kind):
    p = policy_obj
    if p.time_step is None:
        raise _capi.SnbError(f"{type(p).__name__}.time_step is not set (the env sets policy.time_step on reset, "
                             "crowd_sim_plus.py:686-688)")
    return _capi.PolicyCfg(
        policy=kind, max_neighbors=int(getattr(p, "max_neighbors", 10)), time_step=float(p.time_step),
        neighbor_dist=float(getattr(p, "neighbor_dist", 10)), time_horizon=float(getattr(p, "time_horizon", 2.0)),
        time_horizon_obst=float(getattr(p, "time_horizon_obst", 0.5)), policy_radius=float(getattr(p, "radius", 0.3) or 0.3),
        max_speed=float(getattr(p, "max_speed", 1)), safety_space=float(getattr(p, "safety_space", 0.0)),
        sfm_radius=float(getattr(p, "radius", 0.0) or 0.0), A=float(getattr(p, "A", 0.0)), B=float(getattr(p, "B", 1.0)),
        KI=float(getattr(p, "KI", 0.0)), A_static=float(getattr(p, "A_static", 0.0)), B_static=float(getattr(p, "B_static", 1.0)),
        A_bottleneck=float(getattr(p, "A_bottleneck", 0.0)), B_bottleneck=float(getattr(p, "B_bottleneck", 1.0)),
        is_bottleneck=int(bool(getattr(p, "is_bottleneck", False))))


def predict_host(cfg, state, want_neighbors=False):
    """One `policy.predict(state)` through snb_policy_predict_host (host buffers in, ActionXY out)."""
    s = state.self_state
    self8 = (C.c_double * 8)(s.px, s.py, s.vx, s.vy, s.radius, s.gx, s.gy, s.v_pref)
    n = len(state.human_states)
    others = (C.c_double * max(5 * n, 1))()
    for j, o in enumerate(state.human_states):
        others[5 * j:5 * j + 5] = (o.px, o.py, o.vx, o.vy, o.radius)
    segs_l = [c for seg in (state.static_obs or []) for pt in seg for c in pt]
    m = len(segs_l) // 4
    segs = (C.c_double * max(4 * m, 1))(*segs_l)
    out = (C.c_double * 2)()
    nbr = (C.c_int32 * _capi.MAX_AGENTS_PER_ENV)()
    cnt = C.c_int32(0)
    _capi.check(_capi.lib.snb_policy_predict_host(C.byref(cfg), self8, n, others, m, segs, out, nbr, C.byref(cnt)),
                "snb_policy_predict_host")
    action = ActionXY(out[0], out[1])
    if want_neighbors:
        return action, [nbr[k] for k in range(cnt.value)]
    return action


def step_batch(cfg, soa, obstacles=None, want_neighbors=False, stream=None):
    """Policy for all humans of all envs of a CrowdStateSoA: returns v[B,H,2] (+ nbr[B,H,MN], cnt[B,H]) on device."""
    import torch
    B, H, MN = soa.B, soa.H, cfg.max_neighbors
    out = torch.empty(B, H, 2, dtype=torch.float64, device=soa.device)
    nbr = torch.full((B, H, max(MN, 1)), -1, dtype=torch.int32, device=soa.device) if want_neighbors else None
    cnt = torch.zeros(B, H, dtype=torch.int32, device=soa.device) if want_neighbors else None
    status = torch.zeros(1, dtype=torch.int32, device=soa.device)
    st = soa.cstruct()
    _capi.check(_capi.lib.snb_policy_step(C.byref(cfg), C.byref(st), obstacles.handle if obstacles is not None else None,
                                          _capi.ptr(out), _capi.ptr(nbr), _capi.ptr(cnt), _capi.ptr(status),
                                          _capi.stream_ptr(stream)), "snb_policy_step")
    if want_neighbors:
        return out, nbr, cnt, status
    return out, status
