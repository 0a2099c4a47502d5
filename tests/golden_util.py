"""Helpers shared by the CPU (oracle) and GPU (CUDA path) golden-rollout tests."""
import configparser
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HCOLS = ("px", "py", "vx", "vy", "theta", "gx", "gy", "fgx", "fgy", "vpref", "radius")
RCOLS = ("rpx", "rpy", "rvx", "rvy", "rtheta", "rgx", "rgy")
DOOR_SIMS = ("hallway_static", "hallway_static_with_back", "hallway_bottleneck")


def rollout_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "rollout_*.npz")))


def load_rollout(path):
    g = dict(np.load(path, allow_pickle=False))
    g["name"] = os.path.basename(path)[len("rollout_"):-4]
    for k in ("human_policy", "sim"):
        g[k] = str(g[k])
    return g


def door_params(g):
    """(enabled, door_y_mid_min, door_y_mid_max, door_x_mid, door_y_min, door_y_max, door_width)"""
    enabled = g["sim"] in DOOR_SIMS and len(g["segs"]) > 0
    d = np.nan_to_num(g["door"], nan=0.0)
    return (int(enabled),) + tuple(float(x) for x in d)


def human_policy_config(g):
    """the [env] / [humans] keys the human policies read in configure (orca_plus.py:15-27, social_force.py:21-36); values of the
    reference's sicnav_diffusion/configs/env.config as the golden generator used them"""
    cfg = configparser.RawConfigParser()
    cfg.read_dict({"env": {"time_step": str(float(g["time_step"]))},
                   "humans": {"radius": str(float(g["policy_radius"])), "safety_space": str(float(g["safety_space"])), "A": "3.0", "B": "0.18",
                              "KI": "1.0", "A_static": "2.0", "B_static": "0.025", "A_bottleneck": "6.0", "B_bottleneck": "0.12"}})
    return cfg
