"""oracle/scenario_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Host restatement (numpy) of the reference's seeded scenario generators, the checker of csrc/scene_kernels.cu:
CrowdSimPlus.reset -> generate_random_human_position / generate_circle_crossing_human / generate_hallway_human and Human.get_g_xy
(crowd_sim_plus/envs/crowd_sim_plus.py:425-481, 522-605, 658-682; envs/utils/human_plus.py:19-52), consuming the PCG64 stream of
`np.random.default_rng(offset + case)` in the reference's draw order, so scene b is the very scene the reference builds for test case
b.  Pinned: tests/test_capi_cpu.py replays the reference-generated golden episodes' initial states (h0) bit for bit.
Only the constant wall-layout table and the parameter record are shared with the product (snb.scenario).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import math
import os
import sys

import numpy as np
from numpy.linalg import norm

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "safe-interactive-crowdnav_b200"))
from snb.scenario import DOOR_RULES, HALLWAY_RULES, SceneParams, static_obstacles  # noqa: E402,F401  (constant layout table)


def _point_to_segment_dist(x1, y1, x2, y2, x3, y3):
    # utils_plus.point_to_segment_dist (utils_plus.py:73-95)
    px, py = x2 - x1, y2 - y1
    if px == 0 and py == 0:
        return norm((x3 - x1, y3 - y1))
    u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py)
    u = 1 if u > 1 else (0 if u < 0 else u)
    return norm((x1 + u * px - x3, y1 + u * py - y3))


def door_goal(rule, door, n_seg, px, py, fgx, fgy):
    """Human.get_g_xy (human_plus.py:19-52)."""
    if n_seg > 0 and rule in DOOR_RULES and door is not None:
        ys = [py, fgy]
        if np.min(ys) < door["door_y_mid_min"] and np.max(ys) > door["door_y_mid_max"]:
            int_gx, int_gy = door["door_x_mid"], 0.5 * (door["door_y_min"] + door["door_y_max"])
            if np.linalg.norm(np.array([int_gx - px, int_gy - py])) <= door["door_width"] / 2.0:
                return fgx, fgy
            return int_gx, int_gy
    return fgx, fgy


def generate_scene(rule, H, case, phase, p, case_capacity=None):
    """One environment: returns dict(segs, door, humans[H,8]=px,py,gx,gy,fgx,fgy,v_pref,theta, robot=(px,py,gx,gy,theta))."""
    cap = case_capacity or {"val": 1000, "test": 1000}
    offset = {"train": cap["val"] + cap["test"], "val": 0, "test": cap["val"]}[phase]
    segs, door = static_obstacles(rule, p)
    robot = (0.0, -p.circle_radius, 0.0, p.circle_radius, math.pi / 2)
    if case == -1:  # debug layout (crowd_sim_plus.py:676-682)
        assert phase == "test"
        raw = [(0, -6, 0, 5, p.human_v_pref, math.pi / 2), (-5, -5, -5, 5, p.human_v_pref, math.pi / 2),
               (5, -5, 5, 5, p.human_v_pref, math.pi / 2)]
    else:
        rng = np.random.default_rng(offset + case)
        raw = _generate_with_goals(rule, H, rng, p, segs, door)
    humans = np.zeros((len(raw), 8))
    for i, (px, py, fgx, fgy, v_pref, theta) in enumerate(raw):
        gx, gy = door_goal(rule, door, len(segs), px, py, fgx, fgy)
        humans[i] = (px, py, gx, gy, fgx, fgy, v_pref, theta)
    return dict(segs=segs, door=door, humans=humans, robot=robot)


def _generate_with_goals(rule, H, rng, p, segs, door):
    """generate_random_human_position (:425-451), tracking each placed human's CURRENT goal (gx,gy after Human.set -> set_g_xy), which the
    collision tests of later humans read (`agent.gx`, crowd_sim_plus.py:475, :590)."""
    robot = (0.0, -p.circle_radius, 0.0, p.circle_radius, p.robot_radius)
    placed = []
    out = []
    for _ in range(H):
        one, _ = _generate_one(rule, rng, p, segs, robot, placed)
        px, py, fgx, fgy, v_pref, theta = one
        gx, gy = door_goal(rule, door, len(segs), px, py, fgx, fgy)
        placed.append((px, py, gx, gy, p.human_radius))
        out.append(one)
    return out


def _generate_one(rule, rng, p, segs, robot, placed):
    """generate_circle_crossing_human (:454-481) / generate_hallway_human (:522-605) for ONE human;
    `placed` = (px,py,gx,gy,radius) of the humans already in the scene, `robot` likewise."""
    v_pref = p.human_v_pref
    r = p.human_radius
    agents = [robot] + placed
    if rule == "circle_crossing":
        if p.randomize_attributes:
            v_pref = rng.uniform(0.5, 1.5)
        while True:
            angle = rng.random() * np.pi * 2
            px_noise = (rng.random() - 0.5) * v_pref
            py_noise = (rng.random() - 0.5) * v_pref
            px = p.circle_radius * np.cos(angle) + px_noise
            py = p.circle_radius * np.sin(angle) + py_noise
            collide = False
            for (apx, apy, agx, agy, ar) in agents:
                min_dist = r + ar + p.discomfort_dist
                if norm((px - apx, py - apy)) < min_dist or norm((px - agx, py - agy)) < min_dist:
                    collide = True
                    break
            if not collide:
                break
        return (px, py, -px, -py, v_pref, 0), None
    if rule in HALLWAY_RULES:
        eff_h = p.rect_height
        while True:
            if p.randomize_attributes:
                v_pref = rng.uniform(0.5, 1.5)
            dir_sign = 1 if rng.random() < 0.15 else -1
            prob_right = 0.8
            right_num = prob_right if dir_sign > 0 else 1 - prob_right
            wor_sign = -1 if rng.random() < right_num else 1
            prob_cross = 0.3
            if rng.random() < right_num:
                prob_cross = 1 - prob_cross
            cross_sign = -wor_sign if rng.random() < prob_cross else wor_sign
            px = (rng.random()) * 0.5 * wor_sign * (p.rect_width - r * 2)
            py = (rng.random()) * 0.25 * dir_sign * p.circle_radius * (eff_h - r * 2)
            collide = norm((px - robot[0], py - robot[1])) < r + robot[4] + p.discomfort_dist
            if not collide:
                for (apx, apy, agx, agy, ar) in agents:
                    if norm((px - apx, py - apy)) < r + ar:
                        collide = True
                        break
            if not collide:
                for L in segs:
                    if np.abs(_point_to_segment_dist(L[0], L[1], L[2], L[3], px, py)) < (r + 0.01):
                        collide = True
                        break
            if collide:
                eff_h *= 1.1
                continue
            gx = (rng.random()) * 0.5 * cross_sign * (p.rect_width - r * 2)
            gy = (rng.random()) * 0.5 * -dir_sign * p.circle_radius * (eff_h - r * 2)
            collide = False
            for (apx, apy, agx, agy, ar) in agents:
                if norm((gx - agx, gy - agy)) < r + ar:
                    collide = True
                    break
            if not collide:
                for L in segs:
                    if np.abs(_point_to_segment_dist(L[0], L[1], L[2], L[3], gx, gy)) < r:
                        collide = True
                        break
            if not collide:
                break
            eff_h *= 1.1
        return (px, py, gx, gy, v_pref, np.arctan2(gy - py, gx - px)), None
    raise ValueError("Rule doesn't exist (square_crossing is broken in the reference, quirk q9)")
