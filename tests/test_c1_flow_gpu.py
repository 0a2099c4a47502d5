"""GPU: BASELINE configs[0] -- the reference's object-at-a-time plumbing (simple_test.py:109-269 -> Human.act -> policy.predict, one
call per human per step) driven through the DROP-IN policy objects `snb.policy.policy_factory['orca' | 'orca_plus' | 'sfm']`, i.e. the
B = 1 plugin path (JointState -> snb_policy_predict_host -> ActionXY).  The loop is oracle/c1_flow.py's restatement of the reference's
(the real CrowdSimPlus cannot be present next to a GPU, see that file); it must reproduce reference-GENERATED golden episodes."""
import numpy as np
import pytest

import c1_flow
from golden_util import human_policy_config as _env_config, load_rollout, rollout_files

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", rollout_files(), ids=lambda p: p.split("rollout_")[-1][:-4])
def test_reference_plumbing_with_drop_in_policies_replays_golden_episode(path):
    from snb.policy.policy_factory import policy_factory
    from snb.utils.state_plus import FullState, JointState, ObservableState
    g = load_rollout(path)
    tol = 1e-9 if g["human_policy"] != "sfm" else 1e-8
    n = 0
    for k, (hs, rs) in enumerate(c1_flow.run_episode(g, policy_factory, (FullState, ObservableState, JointState), _env_config(g))):
        ref_h = g["H_states"][k][:, :7]
        assert np.max(np.abs(hs - ref_h)) < tol, (k, np.max(np.abs(hs - ref_h)))
        assert np.max(np.abs(rs - g["R_states"][k])) < tol, k
        n += 1
    assert n == len(g["actions"]) >= 8


def test_policy_factory_objects_are_the_reference_surface():
    from snb.policy.policy_factory import policy_factory
    for name in ("orca", "orca_plus", "sfm", "linear"):
        p = policy_factory[name]()
        for attr in ("trainable", "phase", "model", "device", "last_state", "time_step", "env", "kinematics"):
            assert hasattr(p, attr), (name, attr)
        assert p.kinematics == "holonomic" and callable(p.predict) and callable(p.configure)
