"""oracle/ref_shims.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Import helpers that make parts of the Python reference under /root/reference importable
in the BUILD CONTAINER ONLY (the GPU box has no /root/reference).  Used by
oracle/gen_golden.py to generate tests/golden/* and by tests that are skipped when the
reference is absent.  Recipes: SURVEY.md Appendix D.
"""
import os
import pickle
import sys
import types

REF = os.environ.get("SNB_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def have_reference():
    return os.path.isdir(os.path.join(REF, "crowd_sim_plus"))


def install_crowd_sim_shims():
    """Namespace shims so crowd_sim_plus.* imports without gym / matplotlib / the real rvo2."""
    import torch  # noqa: F401  (must be imported before the matplotlib stubs exist)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    shim = os.path.join(_HERE, "rvo2_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    for name, rel in (("crowd_sim_plus", "crowd_sim_plus"), ("crowd_sim_plus.envs", "crowd_sim_plus/envs"),
                      ("crowd_sim_plus.envs.policy", "crowd_sim_plus/envs/policy"),
                      ("crowd_sim_plus.envs.utils", "crowd_sim_plus/envs/utils")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, rel)]
            sys.modules[name] = m
    # stubs for modules crowd_sim_plus.py imports at top level but the step path never uses
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Env:  # gym.Env
        pass

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    gym = stub("gym", Env=_Env)
    gym.spaces = stub("gym.spaces", Discrete=_Any, Box=_Any, Dict=_Any)
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.patches", "matplotlib.animation",
              "matplotlib.collections", "matplotlib.cm", "matplotlib.colors", "matplotlib.transforms"):
        m = stub(n)

        def _ga(k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()
        m.__getattr__ = _ga  # type: ignore
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].lines = sys.modules["matplotlib.lines"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]


def load_jmid_reference(ckpt="sim_gen_sicnav_p_midjp_cvg_epoch121.pt", diffnet="JointPredictionTransformerConcatLinear"):
    """Returns (DiffusionTraj module with the shipped weights, raw checkpoint dict)."""
    import torch
    import torch.nn as nn
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from sicnav_diffusion.JMID.MID.models import diffusion as D

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except Exception:
                return type(name, (nn.Module,), {})

    class P:
        Unpickler = U
        load = pickle.load
        __name__ = "pickle"

    path = os.path.join(REF, "sicnav_diffusion/JMID/MID/checkpoints/sim_inference_checkpoints", ckpt)
    ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=P)
    net = getattr(D, diffnet)(2, 256, 3, False)
    dt = D.DiffusionTraj(net, D.VarianceSchedule(num_steps=100, beta_T=5e-2, mode="linear"))
    dt.load_state_dict({k[len("vel_predictor."):]: v for k, v in ck["ddpm"].items() if k.startswith("vel_predictor.")})
    dt.eval()
    return dt, ck
