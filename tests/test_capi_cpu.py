"""CPU tests of the boundary: libsnb.so loads, exports every symbol include/snb.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "snb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from snb import _capi
    syms = _declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(_capi.lib, s)]
    assert not missing, missing
    assert _capi.lib.snb_version() == 100


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (checked against sizes computed from the header's field lists)."""
    from snb import _capi
    assert C.sizeof(_capi.PolicyCfg) == 8 + 8 * 15 + 8
    assert C.sizeof(_capi.DoorCfg) == 8 + 6 * 8
    assert C.sizeof(_capi.RewardCfg) == 5 * 8 + 8 + 4 * 8
    assert C.sizeof(_capi.CrowdState) == 16 + 22 * 8 + 8
    assert C.sizeof(_capi.JmidWeights) == (4 * 5 + 3 * 12 + 3) * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from snb import _capi
    from snb.policy import ORCA
    from snb.state import CrowdStateSoA
    from snb.utils.state_plus import FullState, JointState
    with pytest.raises(_capi.SnbError):
        CrowdStateSoA(4, 3, 1, "cpu")
    pol = ORCA()
    pol.time_step = 0.25
    with pytest.raises(_capi.SnbError) as ei:
        pol.predict(JointState(FullState(0, 0, 0, 0, 0.3, 1, 1, 1.0, 0.0), []))
    assert "failed with code -2" in str(ei.value)      # SNB_ECUDA: no device, nothing computed on the host


def test_policy_surface_matches_reference_names():
    from snb.policy import ORCA, ORCAPlus, SFM, Policy, policy_factory
    assert set(policy_factory) >= {"none", "linear", "orca", "orca_plus", "sfm"}
    o = ORCA()
    for attr, val in dict(name='ORCA', trainable=False, multiagent_training=None, kinematics='holonomic', safety_space=0,
                          neighbor_dist=10, max_neighbors=10, time_horizon=2.0, time_horizon_obst=0.5, radius=0.3,
                          max_speed=1, sim=None).items():
        assert getattr(o, attr) == val, attr
    assert isinstance(ORCAPlus(), ORCA) and isinstance(SFM(), Policy)
    assert SFM().name == 'sfm' and SFM().is_bottleneck is False
    for m in ("configure", "set_phase", "set_device", "set_env", "get_model", "predict", "reach_destination"):
        assert hasattr(Policy, m)


def test_scenario_generator_reproduces_reference_scenes():
    """snb.scenario (host-side seeded reset) against the initial states of the reference-generated episodes."""
    import numpy as np
    from golden_util import load_rollout, rollout_files
    import scenario_oracle as scenario
    n = 0
    for f in rollout_files():
        g = load_rollout(f)
        if int(g["starts_moving"]) != 0:
            continue
        segs = g["segs"]
        width = 1.75 if not len(segs) else (2.0 if g["sim"] == "hallway_squeeze" else float(2 * abs(segs[0][0])))
        p = scenario.SceneParams(float(g["circle_radius"]), width, 4.0, float(g["h0"][0, 10]), 1.5, float(g["robot_radius"]), 0.2, True)
        sc = scenario.generate_scene(g["sim"], int(g["H"]), int(g["case"]), "test", p)
        assert np.array_equal(sc["humans"], g["h0"][:, [0, 1, 5, 6, 7, 8, 9, 4]]), g["name"]
        if len(segs):
            assert np.array_equal(sc["segs"], segs)
        n += 1
    assert n >= 8


def test_predictor_surface_matches_reference_names():
    """The inner predictor entry keeps the reference's parameter names and order (models/diffusion.py:478-491), the wrapper class its
    method names (mid_sim_wrapper.py:172-509); unsupported modes are refused before any device work."""
    import inspect
    from snb import _capi
    from snb.jmid import DiffusionTraj
    from snb.jmid.forecaster import HumanTrajectoryForecasterSim
    params = list(inspect.signature(DiffusionTraj.sample_sicnav_inference).parameters)
    assert params[:12] == ["self", "num_points", "context", "sample", "bestof", "point_dim", "flexibility", "ret_traj", "sampling", "step",
                           "with_constraints", "dynamics"]
    sig = inspect.signature(DiffusionTraj.sample_sicnav_inference)
    assert sig.parameters["sampling"].default == "ddpm" and sig.parameters["step"].default == 100      # the reference's defaults
    for name in ("update_state_hists", "predict_ret_best", "predict", "get_most_likely_samples"):
        assert callable(getattr(HumanTrajectoryForecasterSim, name))
    assert list(inspect.signature(HumanTrajectoryForecasterSim.__init__).parameters)[1:3] == ["env_config", "mid_config_file"]
    d = DiffusionTraj({}, joint=True)
    import torch
    ctx = torch.zeros(2, 256)
    for bad in (dict(sampling="ddpm"), dict(sampling="ddim", ret_traj=True), dict(sampling="ddim", point_dim=3), dict(sampling="ddim", step=30),
                dict(sampling="ddim", step=0)):
        with pytest.raises(_capi.SnbError):
            d.sample_sicnav_inference(8, ctx, 4, True, **bad)
    with pytest.raises(_capi.SnbError):
        d.sample_sicnav_inference(8, torch.zeros(2, 128), 4, True, sampling="ddim", step=20)
