"""ORCAPlus: ORCA + static line-segment obstacles, drop-in for crowd_sim_plus/envs/policy/orca_plus.py."""
import logging

from .. import _capi
from .orca import ORCA


class ORCAPlus(ORCA):
    _KIND = _capi.POLICY_ORCA_PLUS

    def __init__(self):
        super().__init__()

    def configure(self, config, section='orca_plus'):
        try:
            self.time_step = config.getfloat('env', 'time_step')
        except Exception:
            logging.warning("[ORCA_PLUS POLICY] problem with policy config")
        self.radius = config.getfloat(section, 'radius')
        self.safety_space = config.getfloat(section, 'safety_space')
        return
