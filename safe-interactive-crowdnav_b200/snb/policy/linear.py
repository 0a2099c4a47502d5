"""Straight-to-goal policy (crowd_sim_plus/envs/policy/linear.py:6-23); host-only, used as a stand-in robot."""
import math

from ..utils.action import ActionXY
from .policy import Policy


class Linear(Policy):
    def __init__(self):
        super().__init__()
        self.trainable = False
        self.kinematics = 'holonomic'
        self.multiagent_training = True

    def configure(self, config):
        return

    def predict(self, state):
        s = state.self_state
        theta = math.atan2(s.gy - s.py, s.gx - s.px)
        return ActionXY(math.cos(theta) * s.v_pref, math.sin(theta) * s.v_pref)
