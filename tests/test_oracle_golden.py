"""CPU tests: the C oracle (oracle/crowd_oracle.c) against fixtures produced by the Python reference
(oracle/gen_golden.py).  These pin the oracle before any GPU parity claim is made against it."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from golden_util import GOLDEN, HCOLS, RCOLS, door_params, load_rollout, rollout_files


def test_sfm_cases_match_reference():
    g = np.load(f"{GOLDEN}/sfm_cases.npz")
    L = ol.lib()
    worst = 0.0
    for i in range(int(g["n"])):
        cfg = ol.default_policy_cfg("sfm", is_bottleneck=int(g["bottleneck"][i]))
        self8 = np.ascontiguousarray(g[f"self8_{i}"], np.float64)
        others = np.ascontiguousarray(g[f"others_{i}"], np.float64).ravel()
        segs = np.ascontiguousarray(g[f"segs_{i}"], np.float64).ravel()
        out = np.zeros(2)
        L.orc_sfm_predict(C.byref(cfg), ol.dptr(self8), len(others) // 5, ol.dptr(others) if len(others) else None,
                          len(segs) // 4, ol.dptr(segs) if len(segs) else None, ol.dptr(out))
        ref = g["out"][i]
        worst = max(worst, float(np.max(np.abs(out - ref) / np.maximum(1.0, np.abs(ref)))))
    assert worst < 1e-12, worst


def test_sfm_survey_known_answer():
    # SURVEY.md Appendix C.3 (reference social_force.py with the shipped env.config)
    L = ol.lib()
    self8 = np.array([.1, -.2, .3, .4, .2, 1.0, 3.0, 1.2])
    others = np.array([[.6, .1, -.2, 0, .2], [-.5, -.9, .1, .5, .25]]).ravel()
    segs = np.array([[-.875, -4, -.875, 4], [.875, -4, .875, 4]], np.float64).ravel()
    out = np.zeros(2)
    cfg = ol.default_policy_cfg("sfm")
    L.orc_sfm_predict(C.byref(cfg), ol.dptr(self8), 2, ol.dptr(others), 2, ol.dptr(segs), ol.dptr(out))
    assert np.allclose(out, [0.10987619464893483, 0.4990339500221377], rtol=0, atol=1e-14)
    segs3 = np.concatenate([segs, [-.875, 0, -.5, 0]])
    cfg = ol.default_policy_cfg("sfm", is_bottleneck=1)
    L.orc_sfm_predict(C.byref(cfg), ol.dptr(self8), 2, ol.dptr(others), 3, ol.dptr(segs3), ol.dptr(out))
    assert np.allclose(out, [0.15197744173784936, 0.48500020099249946], rtol=0, atol=1e-14)


def test_segment_geometry_matches_reference():
    g = np.load(f"{GOLDEN}/geometry_cases.npz")
    L = ol.lib()
    out = np.zeros(5)
    worst = 0.0
    nskip = 0
    for inp, ref in zip(g["segseg_in"], g["segseg_out"]):
        a = np.ascontiguousarray(inp, np.float64)
        # Near-parallel segments whose unit-vector cross product is a rounding residue (0 < |cz| < 1e-9):
        # the reference then divides LAPACK-LU determinants of a numerically singular 3x3 by denom ~ 1e-32
        # (utils_plus.py:300-306) and its own answer is rounding noise.  Exactly parallel (cz == 0, e.g. travel
        # along an axis-aligned wall) and all generic cases are checked.
        A3, B3 = np.array([*(a[2:4] - a[0:2]), 0.0]), np.array([*(a[6:8] - a[4:6]), 0.0])
        if np.linalg.norm(A3) > 1e-8 and np.linalg.norm(B3) > 1e-8:
            cz = np.cross(A3 / np.linalg.norm(A3), B3 / np.linalg.norm(B3))[2]
            if 0 < abs(cz) < 1e-9:
                nskip += 1
                continue
        L.orc_closest_distance_between_line_segments(ol.dptr(a[0:2].copy()), ol.dptr(a[2:4].copy()), ol.dptr(a[4:6].copy()),
                                                     ol.dptr(a[6:8].copy()), ol.dptr(out))
        worst = max(worst, float(np.max(np.abs(out - ref))))
    assert worst < 1e-12, worst
    assert nskip < 0.35 * len(g["segseg_in"])  # the synthetic parallel constructions


def test_action_clamp_matches_reference():
    g = np.load(f"{GOLDEN}/geometry_cases.npz")
    L = ol.lib()
    out = np.zeros(2)
    worst = 0.0
    nclamped = 0
    for inp, ref, li in zip(g["clamp_in"], g["clamp_out"], g["clamp_layout"]):
        segs = np.ascontiguousarray(g[f"layout_{li}"], np.float64).ravel()
        pose = np.array(inp[0:3])
        act = np.array(inp[6:8])
        L.orc_constrain_action(ol.dptr(pose), float(inp[3]), float(inp[4]), int(inp[5]), ol.dptr(act), len(segs) // 4,
                               ol.dptr(segs), ol.dptr(out))
        worst = max(worst, float(np.max(np.abs(out - ref))))
        nclamped += int(np.any(out != act))
    assert worst < 1e-9, worst
    assert nclamped > 100


def _env_from_golden(g):
    H = int(g["H"])
    env = ol.EnvArrays(1, H, g["segs"], rradius=float(g["robot_radius"]), rvpref=float(g["robot_vpref"]),
                       robot_kinematics=ol.KIN_UNICYCLE if bool(g["unicycle"]) else ol.KIN_HOLONOMIC)
    for j, n in enumerate(HCOLS):
        getattr(env, n)[:] = g["h0"][:, j]
    for j, n in enumerate(RCOLS):
        getattr(env, n)[:] = g["r0"][j]
    env.human_time[:] = g["human_times0"]
    env.global_time[:] = float(g["global_time0"])
    pcfg = ol.default_policy_cfg(g["human_policy"], time_step=float(g["time_step"]), safety_space=float(g["safety_space"]),
                                 sfm_radius=float(g["policy_radius"]), is_bottleneck=int(g["is_bottleneck"]))
    door = ol.DoorCfg(*door_params(g))
    rcfg = ol.default_reward_cfg(time_limit=float(g["time_limit"]))
    return env, pcfg, door, rcfg


@pytest.mark.parametrize("path", rollout_files(), ids=lambda p: p.split("rollout_")[-1][:-4])
def test_rollout_matches_reference_env(path):
    """Step-by-step replay of a reference CrowdSimPlus episode (humans on orca.py/orca_plus.py/social_force.py,
    reference clamp, reference reward/info) through oracle orc_env_step."""
    g = load_rollout(path)
    env, pcfg, door, rcfg = _env_from_golden(g)
    tol = 1e-9
    for k in range(len(g["reward"])):
        reward, dmin, flags = ol.env_step(pcfg, door, rcfg, env, g["actions"][k])
        hs = np.stack([getattr(env, n) for n in HCOLS], 1)
        rs = np.array([getattr(env, n)[0] for n in RCOLS])
        assert np.max(np.abs(hs - g["H_states"][k])) < tol, (k, np.max(np.abs(hs - g["H_states"][k])))
        assert np.max(np.abs(rs - g["R_states"][k])) < tol, k
        assert abs(reward[0] - g["reward"][k]) < 1e-9, (k, reward[0], g["reward"][k])
        assert int(flags[0]) == int(g["flags"][k]), (k, int(flags[0]), int(g["flags"][k]))
        if not np.isnan(g["dmin"][k]):
            assert abs(dmin[0] - g["dmin"][k]) < 1e-9
        assert np.max(np.abs(env.human_time - g["human_times"][k])) < 1e-9, k
        assert abs(env.global_time[0] - g["global_time"][k]) < 1e-12
