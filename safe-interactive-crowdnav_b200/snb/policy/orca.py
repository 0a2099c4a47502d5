"""ORCA human policy, drop-in for crowd_sim_plus/envs/policy/orca.py (same class name, attributes, defaults and
`predict(state) -> ActionXY`).  Where the reference builds a throw-away rvo2.PyRVOSimulator per call
(orca.py:94-129), this class hands the JointState to the sm_100a ORCA kernel through the C ABI; `predict_batch`
runs every human of every environment of a CrowdStateSoA in one launch."""
from .. import _capi
from . import _device_policy as _dp
from .policy import Policy


class ORCA(Policy):
    _KIND = _capi.POLICY_ORCA

    def __init__(self):
        super().__init__()
        self.name = 'ORCA'
        self.trainable = False
        self.multiagent_training = None
        self.kinematics = 'holonomic'
        # constants of orca.py:59-66 (quirk q8: hard-coded, not read from the config)
        self.safety_space = 0
        self.neighbor_dist = 10
        self.max_neighbors = 10
        self.time_horizon = 2.0
        self.time_horizon_obst = 0.50
        self.radius = 0.3
        self.max_speed = 1
        self.sim = None          # kept for attribute compatibility; no simulator object exists here
        self.last_neighbors = None

    def configure(self, config):
        # one-argument signature like the reference: Human.__init__ calls configure(config, section) inside
        # try/except, so plain 'orca' humans keep the defaults above (quirk q6)
        return

    def set_phase(self, phase):
        return

    def _cfg(self):
        return _dp.policy_cfg(self, self._KIND)

    def predict(self, state):
        action, nbrs = _dp.predict_host(self._cfg(), state, want_neighbors=True)
        self.last_state = state
        self.last_neighbors = nbrs
        return action

    def predict_batch(self, soa, obstacles=None, want_neighbors=False, stream=None):
        return _dp.step_batch(self._cfg(), soa, obstacles, want_neighbors, stream)
