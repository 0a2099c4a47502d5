"""Seeded scenario generation for the batched simulator: the synthetic-input generator of every benchmark config.

Restates CrowdSimPlus.reset / generate_static_obstacles / generate_circle_crossing_human / generate_hallway_human
(crowd_sim_plus/envs/crowd_sim_plus.py:609-764, 322-422, 454-481, 522-605) on the host with numpy, consuming the
PCG64 stream of `np.random.default_rng(offset + case)` in the reference's draw order, so env b of a batch is the
very scene the reference builds for test case b.  (Runs once per episode; the per-step hot path is on the GPU.)
"""
import math

import numpy as np
from numpy.linalg import norm

DOOR_RULES = ("hallway_static", "hallway_static_with_back", "hallway_bottleneck")
HALLWAY_RULES = ("hallway", "hallway_static", "hallway_bottleneck", "hallway_squeeze", "rectangle",
                 "hallway_static_with_back", "left_wall", "no_walls")


def _point_to_segment_dist(x1, y1, x2, y2, x3, y3):
    # utils_plus.point_to_segment_dist (utils_plus.py:73-95)
    px, py = x2 - x1, y2 - y1
    if px == 0 and py == 0:
        return norm((x3 - x1, y3 - y1))
    u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py)
    u = 1 if u > 1 else (0 if u < 0 else u)
    return norm((x1 + u * px - x3, y1 + u * py - y3))


class SceneParams:
    """The [sim]/[humans]/[robot] numbers the generators read (CrowdSimPlus.configure, crowd_sim_plus.py:132-166)."""

    def __init__(self, circle_radius, rect_width, rect_height, human_radius, human_v_pref, robot_radius,
                 discomfort_dist, randomize_attributes):
        self.circle_radius, self.rect_width, self.rect_height = circle_radius, rect_width, rect_height
        self.human_radius, self.human_v_pref, self.robot_radius = human_radius, human_v_pref, robot_radius
        self.discomfort_dist, self.randomize_attributes = discomfort_dist, randomize_attributes


def static_obstacles(rule, p):
    """generate_static_obstacles (crowd_sim_plus.py:322-422) -> (segments [n,4], door dict or None)."""
    W, Hh, R = p.rect_width, p.rect_height, p.circle_radius
    door = None
    segs = []
    if rule in ("hallway_static", "hallway_static_with_back", "hallway_bottleneck", "hallway_squeeze"):
        door_y_max = R - p.robot_radius * 2.0
        door_y_min = -R + p.robot_radius * 2.0
        door_x_mid = 0.0
        door_y_mid_max = door_y_max + (door_y_min - door_y_max) * 0.40
        door_y_mid_min = door_y_max + (door_y_min - door_y_max) * (1.0 - 0.40)
        door_width = 0.5 * W if rule == "hallway_squeeze" else 1.0
        xl = door_x_mid - door_width / 2.0
        xlm = xl + ((-W * 0.5) - xl) * 0.75
        xr = door_x_mid + door_width / 2.0
        xrm = xr + (W * 0.5 - xr) * 0.75
        door = dict(door_y_mid_min=door_y_mid_min, door_y_mid_max=door_y_mid_max, door_x_mid=door_x_mid,
                    door_y_min=door_y_min, door_y_max=door_y_max, door_width=door_width)
        if rule == "hallway_squeeze":
            segs = [[(-W * 0.5, -R * 2.5), (xl, 0)], [(xl, 0), (-W * 0.5, R * 2.5)],
                    [(W * 0.5, -R * 2.5), (xr, 0)], [(xr, 0), (W * 0.5, R * 2.5)]]
        else:
            segs = [[(-W * 0.5, -Hh), (-W * 0.5, Hh)], [(W * 0.5, -Hh), (W * 0.5, Hh)]]
            if "hallway_static" in rule:
                segs += [[(-W * 0.5, door_y_min), (xlm, door_y_min)], [(xlm, door_y_min), (xl, door_y_mid_min)],
                         [(xl, door_y_mid_min), (xl, door_y_mid_max)], [(xl, door_y_mid_max), (xlm, door_y_max)],
                         [(xlm, door_y_max), (-W * 0.5, door_y_max)],
                         [(W * 0.5, door_y_min), (xrm, door_y_min)], [(xrm, door_y_min), (xr, door_y_mid_min)],
                         [(xr, door_y_mid_min), (xr, door_y_mid_max)], [(xr, door_y_mid_max), (xrm, door_y_max)],
                         [(xrm, door_y_max), (W * 0.5, door_y_max)]]
            elif rule == "hallway_bottleneck":
                segs += [[(-W * 0.5, 0), (xl, 0)], [(xr, 0), (W * 0.5, 0)]]
            if rule == "hallway_static_with_back":
                segs += [[(-W * 0.5, -Hh * 0.5), (W * 0.5, -Hh * 0.5)], [(-W * 0.5, Hh * 0.5), (W * 0.5, Hh * 0.5)]]
    elif rule == "hallway":
        segs = [[(-W * 0.5, -Hh), (-W * 0.5, Hh)], [(W * 0.5, -Hh), (W * 0.5, Hh)]]
    elif rule == "rectangle":
        segs = [[(-W * 0.5, -Hh * 0.5), (-W * 0.5, Hh * 0.5)], [(W * 0.5, -Hh * 0.5), (W * 0.5, Hh * 0.5)],
                [(-W * 0.5, -Hh * 0.5), (W * 0.5, -Hh * 0.5)], [(-W * 0.5, Hh * 0.5), (W * 0.5, Hh * 0.5)]]
    elif rule == "left_wall":
        segs = [[(-W * 0.5, -Hh * 1000), (-W * 0.5, Hh * 1000)]]
    arr = np.asarray(segs, np.float64).reshape(-1, 4)
    return arr, door


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
