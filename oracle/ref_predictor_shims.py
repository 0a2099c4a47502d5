"""oracle/ref_predictor_shims.py -- TEST INFRASTRUCTURE (build container only).

Makes the reference's full predictor stack (sicnav_diffusion/JMID/mid_sim_wrapper.py -> MID -> Trajectron encoder ->
diffusion) importable on Python 3.12 without ncls / orjson / easydict / tensorboardX / matplotlib, by stubbing exactly
those modules.  Used by oracle/gen_golden.py to record golden vectors of the context encoder and of
HumanTrajectoryForecasterSim.predict_ret_best."""
import collections
import collections.abc
import json
import sys
import types

import ref_shims


def install():
    import torch
    for n in ("Sequence", "Mapping", "MutableMapping", "Iterable"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    ref_shims.install_crowd_sim_shims()

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setattr__(k, v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    class _NCLS:
        def __init__(self, *a, **k):
            pass

    stub("easydict", EasyDict=EasyDict)
    stub("orjson", loads=json.loads, dumps=lambda o, **k: json.dumps(o).encode())
    stub("tensorboardX", SummaryWriter=object)
    stub("ncls", NCLS=_NCLS)
    for n in ("seaborn", "matplotlib.patheffects", "matplotlib.ticker", "matplotlib.gridspec"):
        stub(n)
    if ref_shims.REF not in sys.path:
        sys.path.insert(0, ref_shims.REF)
    # the reference was written for torch 1.13 (torch.load default weights_only=False); its checkpoints pickle nn.Modules
    if not getattr(torch.load, "_snb_patched", False):
        _orig = torch.load

        def _load(*a, **k):
            k.setdefault("weights_only", False)
            return _orig(*a, **k)
        _load._snb_patched = True
        torch.load = _load
    return EasyDict
