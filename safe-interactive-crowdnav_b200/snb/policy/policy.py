"""Plugin base class with the attribute / method surface of crowd_sim_plus/envs/policy/policy.py:5-50."""
import math


class Policy(object):
    def __init__(self):
        self.trainable = False
        self.phase = None
        self.model = None
        self.device = None
        self.last_state = None
        self.time_step = None
        self.env = None
        self.init_weights = False

    def configure(self, config):
        raise NotImplementedError

    def set_phase(self, phase):
        self.phase = phase

    def set_device(self, device):
        self.device = device

    def set_env(self, env):
        self.env = env

    def get_model(self):
        return self.model

    def predict(self, state):
        """state (JointState) -> action"""
        raise NotImplementedError

    @staticmethod
    def reach_destination(state):
        s = state.self_state
        return math.hypot(s.py - s.gy, s.px - s.gx) < s.radius
