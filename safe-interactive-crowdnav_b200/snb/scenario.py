"""Static parts of the scenario set-up of the batched simulator: the rule names, the [sim] / [humans] / [robot] numbers the
generators read and the wall layouts of generate_static_obstacles (crowd_sim_plus/envs/crowd_sim_plus.py:322-422).

The seeded rejection-sampled human placement (generate_circle_crossing_human / generate_hallway_human, :454-481, :522-605) runs on
the device (csrc/scene_kernels.cu, snb_scene_reset); its host restatement is test infrastructure and lives in
oracle/scenario_oracle.py.
"""
import math

import numpy as np

DOOR_RULES = ("hallway_static", "hallway_static_with_back", "hallway_bottleneck")
HALLWAY_RULES = ("hallway", "hallway_static", "hallway_bottleneck", "hallway_squeeze", "rectangle",
                 "hallway_static_with_back", "left_wall", "no_walls")


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


def debug_layout(p):
    """The fixed 3-human layout of test case -1 (crowd_sim_plus.py:676-682): rows px, py, gx, gy, v_pref, theta."""
    return np.array([(0, -6, 0, 5, p.human_v_pref, math.pi / 2), (-5, -5, -5, 5, p.human_v_pref, math.pi / 2),
                     (5, -5, 5, 5, p.human_v_pref, math.pi / 2)], np.float64)
