"""Social-Force human policy, drop-in for crowd_sim_plus/envs/policy/social_force.py (class SFM, name 'sfm')."""
import logging

from .. import _capi
from . import _device_policy as _dp
from .policy import Policy


class SFM(Policy):
    _KIND = _capi.POLICY_SFM

    def __init__(self):
        super().__init__()
        self.name = 'sfm'
        self.trainable = False
        self.multiagent_training = None
        self.kinematics = 'holonomic'
        self.is_bottleneck = False

    def configure(self, config, section='sfm'):
        try:
            self.time_step = config.getfloat('env', 'time_step')
        except Exception:
            logging.warning("[SFM POLICY] problem with policy config")
        for key in ('radius', 'A', 'B', 'KI', 'A_static', 'B_static', 'A_bottleneck', 'B_bottleneck'):
            setattr(self, key, config.getfloat(section, key))
        return

    def _cfg(self):
        return _dp.policy_cfg(self, self._KIND)

    def predict(self, state):
        self.last_state = state
        return _dp.predict_host(self._cfg(), state)

    def predict_batch(self, soa, obstacles=None, stream=None):
        return _dp.step_batch(self._cfg(), soa, obstacles, False, stream)
