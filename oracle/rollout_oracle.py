"""oracle/rollout_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of one test episode of the reference, batched over test cases: CrowdSimPlus.reset (seeded scene + `starts_moving`
warm-up, crowd_sim_plus.py:609-764) and the loop of simple_test.py:216-269 with the Linear robot (envs/policy/linear.py:16-23), on
top of the C step oracle (oracle/crowd_oracle.c, pinned to reference-generated episodes) and the scenario oracle.  Returns the
per-step flag words / dmin and the episode counters simple_test.py:306-319 pickles -- the checker of snb/rollout.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import math

import numpy as np

import oracle_lib as ol
import scenario_oracle as SO


def reset(cases, H, policy="orca", rule="circle_crossing", circle_radius=4.0, rect_width=1.75, rect_height=4.0, human_radius=0.3,
          human_v_pref=1.5, robot_radius=0.25, robot_v_pref=1.0, discomfort_dist=0.2, randomize=True, starts_moving=10, dt=0.25,
          safety_space=0.0, time_limit=30.0, n_threads=4, reward=None):
    p = SO.SceneParams(circle_radius, rect_width, rect_height, human_radius, human_v_pref, robot_radius, discomfort_dist, randomize)
    segs, door = SO.static_obstacles(rule, p)
    B = len(cases)
    env = ol.EnvArrays(B, H, segs if len(segs) else None, rradius=robot_radius, rvpref=robot_v_pref)
    for b, case in enumerate(cases):
        sc = SO.generate_scene(rule, H, int(case), "test", p)
        h = sc["humans"]
        s = slice(b * H, (b + 1) * H)
        env.px[s], env.py[s], env.gx[s], env.gy[s], env.fgx[s], env.fgy[s], env.vpref[s], env.theta[s] = (h[:, k] for k in range(8))
    env.radius[:] = human_radius
    env.rpx[:] = 0.0; env.rpy[:] = -circle_radius; env.rgx[:] = 0.0; env.rgy[:] = circle_radius; env.rtheta[:] = math.pi / 2
    pcfg = ol.default_policy_cfg(policy, time_step=dt, safety_space=safety_space)
    rcfg = ol.default_reward_cfg(time_limit=time_limit, **(reward or {}))
    dcfg = ol.DoorCfg(enabled=0) if (door is None or rule not in SO.DOOR_RULES or not len(segs)) else ol.DoorCfg(enabled=1, **door)
    env.global_time[:] = -starts_moving * dt if starts_moving > 0 else 0.0
    for _ in range(starts_moving):
        ol.env_step(pcfg, dcfg, rcfg, env, np.zeros((B, 2)), n_threads=n_threads)
    env.prev_dist[:] = np.hypot(env.rpx - env.rgx, env.rpy - env.rgy)
    return env, pcfg, dcfg, rcfg


def linear_action(env):
    """Linear.predict (linear.py:16-23) for every environment."""
    theta = np.arctan2(env.rgy - env.rpy, env.rgx - env.rpx)
    return np.stack([np.cos(theta) * env.rvpref, np.sin(theta) * env.rvpref], 1)


def run_episodes(env, pcfg, dcfg, rcfg, dt=0.25, max_steps=None, actions=None, n_threads=4):
    """Steps every environment to done.  actions: optional callable(step, env) -> [B,2] (default: the Linear robot).
    Returns (metrics [B,9] float64 in snb.dist.METRIC_COLUMNS order, list of per-step (flags, dmin, active_before))."""
    B = env.B
    m = np.zeros((B, 9)); m[:, 8] = np.inf
    live = np.ones(B, bool)
    trace = []
    max_steps = max_steps or int(math.ceil(rcfg.time_limit / dt)) + 2
    for k in range(max_steps):
        if not live.any():
            break
        a = linear_action(env) if actions is None else actions(k, env)
        reward, dmin, flags = ol.env_step(pcfg, dcfg, rcfg, env, a, active=live.astype(np.uint8), n_threads=n_threads)
        trace.append((flags.copy(), dmin.copy(), live.copy()))
        f = flags
        m[live, 2] += 1; m[live, 3] += dt
        for col, bit in ((4, 4), (5, 8), (6, 16), (7, 32)):
            m[live & ((f & bit) != 0), col] += 1
        m[live, 8] = np.minimum(m[live, 8], dmin[live])
        m[live & ((f & 1) != 0), 0] = 1; m[live & ((f & 2) != 0), 1] = 1
        live &= (f & 64) == 0
    return m, trace
