"""CPU: the restated reference plumbing loop (oracle/c1_flow.py) itself, pinned to the reference-generated golden episodes with
ORACLE-backed policy objects (C oracle behind the same predict(JointState) -> ActionXY surface).  tests/test_c1_flow_gpu.py then swaps
in the drop-in policies (CUDA kernel, B = 1 plugin path) and asks for the same episodes."""
import collections
import ctypes as C

import numpy as np
import pytest

import c1_flow
import oracle_lib as ol
from golden_util import human_policy_config as _env_config, load_rollout, rollout_files

FullState = collections.namedtuple("FullState", "px py vx vy radius gx gy v_pref theta")
ObservableState = collections.namedtuple("ObservableState", "px py vx vy radius")
JointState = collections.namedtuple("JointState", "self_state human_states static_obs")
ActionXY = collections.namedtuple("ActionXY", "vx vy")


class _OraclePolicy:
    kind = "orca"

    def __init__(self):
        self.time_step = 0.25
        self.is_bottleneck = False
        self.safety_space = 0.0
        self.radius = 0.3

    def configure(self, config, section=None):
        if self.kind == "orca":
            raise TypeError("ORCA.configure takes one argument")       # swallowed by Human.__init__ (quirk q6)
        self.radius = config.getfloat(section, "radius")
        if self.kind == "orca_plus":
            self.safety_space = config.getfloat(section, "safety_space")

    def predict(self, state):
        s = state.self_state
        self8 = np.array([s.px, s.py, s.vx, s.vy, s.radius, s.gx, s.gy, s.v_pref], np.float64)
        others = np.ascontiguousarray([[o.px, o.py, o.vx, o.vy, o.radius] for o in state.human_states], np.float64).reshape(-1, 5)
        segs = np.ascontiguousarray([[a[0], a[1], b[0], b[1]] for a, b in state.static_obs], np.float64).reshape(-1, 4)
        cfg = ol.default_policy_cfg(self.kind, time_step=self.time_step, safety_space=self.safety_space, sfm_radius=self.radius,
                                    is_bottleneck=int(self.is_bottleneck))
        out = np.zeros(2)
        L = ol.lib()
        sp = ol.dptr(segs.reshape(-1)) if len(segs) else None
        if self.kind == "sfm":
            L.orc_sfm_predict(C.byref(cfg), ol.dptr(self8), len(others), ol.dptr(others.reshape(-1)), len(segs), sp, ol.dptr(out))
        else:
            L.orc_orca_predict(C.byref(cfg), ol.dptr(self8), len(others), ol.dptr(others.reshape(-1)), len(segs), sp, ol.dptr(out),
                               None, None, None, None)
        return ActionXY(float(out[0]), float(out[1]))


FACTORY = {k: type(k, (_OraclePolicy,), {"kind": k}) for k in ("orca", "orca_plus", "sfm")}


@pytest.mark.parametrize("path", rollout_files(), ids=lambda p: p.split("rollout_")[-1][:-4])
def test_restated_plumbing_loop_replays_reference_episode(path):
    g = load_rollout(path)
    n = 0
    for k, (hs, rs) in enumerate(c1_flow.run_episode(g, FACTORY, (FullState, ObservableState, JointState), _env_config(g))):
        assert np.max(np.abs(hs - g["H_states"][k][:, :7])) < 1e-9, (k, np.max(np.abs(hs - g["H_states"][k][:, :7])))
        assert np.max(np.abs(rs - g["R_states"][k])) < 1e-9, k
        n += 1
    assert n == len(g["actions"])
