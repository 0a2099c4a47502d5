"""Weight sources of the JMID predictor: reference checkpoints (.pt), their tensor export (.npz) and seeded synthetic sets.

A reference checkpoint is `{"encoder": ModuleDict, "ddpm": state_dict}` (sicnav_diffusion/JMID/MID/mid.py:1231-1232, 1291;
tensor list in SURVEY Appendix B).  Everything here returns the pair the device handles take:
  encoder  {"<module>/<param>": fp32 tensor}   (the flattened `checkpoint["encoder"]`)
  ddpm     {state_dict key without the "vel_predictor." prefix: fp32 tensor}
"""
import math
import os
import pickle

import numpy as np
import torch

ENC_MODULES = {
    "node_history": ("PEDESTRIAN/node_history_encoder", 6),
    "edge_ped": ("PEDESTRIAN->PEDESTRIAN/edge_encoder", 12),
    "edge_robot": ("PEDESTRIAN->JRDB_ROBOT/edge_encoder", 12),
}
ATT = "PEDESTRIAN/edge_influence_encoder"
SHIPPED_NAME = "sim_gen_sicnav_p_midjp_cvg_epoch121"
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


class _TolerantUnpickler(pickle.Unpickler):
    """The reference pickles whole nn.Modules (`registrar.model_dict`, mid.py:1502-1505); their classes live in the reference
    tree.  Classes that cannot be imported are replaced by bare nn.Module subclasses: only the parameters are needed."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return type(name, (torch.nn.Module,), {"__module__": module})


class _TolerantPickle:
    __name__ = "pickle"
    Unpickler = _TolerantUnpickler
    load = staticmethod(lambda f, **kw: _TolerantUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def _strip(ddpm):
    return {(k[len("vel_predictor."):] if k.startswith("vel_predictor.") else k): torch.as_tensor(v) for k, v in ddpm.items()}


def load_checkpoint(path, trusted=True):
    """-> (encoder, ddpm).  `.npz`: the tensor export written by oracle/gen_golden.py ckpt (keys "enc/..." and "ddpm/...", plain
    arrays, no pickle).  Anything else: a reference `.pt`, which pickles nn.Modules, so torch.load runs with weights_only=False --
    that executes code from the file; pass trusted=False to refuse such files."""
    if str(path).endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        enc = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("enc/")}
        ddpm = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ddpm/")}
        return enc, _strip(ddpm)
    if not trusted:
        raise ValueError(f"{path}: a pickled reference checkpoint can only be read with trusted=True (arbitrary code execution)")
    try:
        ck = torch.load(path, map_location="cpu", weights_only=False)
    except (ImportError, AttributeError, ModuleNotFoundError):
        ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
    enc = {}
    md = ck["encoder"]
    items = md.items() if hasattr(md, "items") else md._modules.items()
    for mod_name, mod in items:
        params = mod.named_parameters() if hasattr(mod, "named_parameters") else mod.items()
        for pname, t in params:
            enc[f"{mod_name}/{pname}"] = t.detach() if hasattr(t, "detach") else torch.as_tensor(t)
    return enc, _strip(ck["ddpm"])


def shipped_checkpoint_path():
    """Where the shipped JMID checkpoint can be found on this machine, or None: $SNB_JMID_CHECKPOINT, the repository's tensor
    export (tests/golden/ckpt_jmid_epoch121.npz), or the reference tree's own .pt."""
    cands = [os.environ.get("SNB_JMID_CHECKPOINT"),
             os.path.join(_ROOT, "tests", "golden", "ckpt_jmid_epoch121.npz"),
             os.path.join(os.environ.get("SNB_REFERENCE", "/root/reference"),
                          "sicnav_diffusion/JMID/MID/checkpoints/sim_inference_checkpoints", SHIPPED_NAME + ".pt")]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    return None


def resolve_model_path(model_path):
    """The yaml's `model_path` is relative to the reference root (test_time_configs/mid_jp.yaml:11).  Tries it as given, under
    $SNB_REFERENCE, and finally -- when it names the shipped JMID checkpoint -- the repository's .npz export of it."""
    cands = [model_path, os.path.join(os.environ.get("SNB_REFERENCE", "/root/reference"), model_path)]
    for c in cands:
        if os.path.isfile(c):
            return c
    if SHIPPED_NAME in os.path.basename(model_path):
        p = shipped_checkpoint_path()
        if p:
            return p
    raise FileNotFoundError(f"JMID checkpoint {model_path!r} not found (tried {cands})")


# ---- seeded synthetic weights (random-init of the same architecture; reproducible on any box) ----
def _ddpm_shapes():
    sh = {}

    def csl(name, din, dout):
        sh[f"net.{name}._layer.weight"] = (dout, din)
        sh[f"net.{name}._layer.bias"] = (dout,)
        sh[f"net.{name}._hyper_bias.weight"] = (dout, 259)
        sh[f"net.{name}._hyper_gate.weight"] = (dout, 259)
        sh[f"net.{name}._hyper_gate.bias"] = (dout,)

    csl("concat1", 2, 512)
    for l in range(3):
        p = f"net.transformer_encoder.layers.{l}."
        for n, s in (("self_attn.in_proj_weight", (1536, 512)), ("self_attn.in_proj_bias", (1536,)),
                     ("self_attn.out_proj.weight", (512, 512)), ("self_attn.out_proj.bias", (512,)),
                     ("linear1.weight", (1024, 512)), ("linear1.bias", (1024,)), ("linear2.weight", (512, 1024)),
                     ("linear2.bias", (512,)), ("norm1.weight", (512,)), ("norm1.bias", (512,)), ("norm2.weight", (512,)),
                     ("norm2.bias", (512,))):
            sh[p + n] = s
    csl("concat3", 512, 256)
    csl("concat4", 256, 128)
    csl("linear", 128, 2)
    return sh


def synthetic_ddpm(seed=5):
    """Uniform(+-1/sqrt(fan_in)) like nn.Linear's default, LayerNorm weight 1 +- 0.1; numpy PCG64 stream of `seed`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    for k, s in _ddpm_shapes().items():
        if ".norm" in k:
            v = (1.0 + 0.1 * rng.standard_normal(s)) if k.endswith("weight") else 0.05 * rng.standard_normal(s)
        else:
            b = 1.0 / math.sqrt(s[1]) if len(s) == 2 else 0.05
            v = rng.uniform(-b, b, s)
        w[k] = torch.from_numpy(np.asarray(v, np.float32))
    return w


def synthetic_encoder(seed=9):
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    b = 1.0 / math.sqrt(128)
    for mod, din in ENC_MODULES.values():
        w[f"{mod}/weight_ih_l0"] = rng.uniform(-b, b, (512, din))
        w[f"{mod}/weight_hh_l0"] = rng.uniform(-b, b, (512, 128))
        w[f"{mod}/bias_ih_l0"] = rng.uniform(-b, b, (512,))
        w[f"{mod}/bias_hh_l0"] = rng.uniform(-b, b, (512,))
    for n, s in (("w1.weight", (128, 128)), ("w2.weight", (128, 128)), ("v.weight", (1, 128))):
        w[f"{ATT}/{n}"] = rng.uniform(-b, b, s)
    return {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in w.items()}


def default_weights(allow_synthetic=True):
    """-> (encoder, ddpm, description): the shipped checkpoint when it is on this machine, else the seeded synthetic set."""
    p = shipped_checkpoint_path()
    if p:
        enc, ddpm = load_checkpoint(p)
        return enc, ddpm, f"shipped checkpoint {SHIPPED_NAME} ({os.path.basename(p)})"
    if not allow_synthetic:
        raise FileNotFoundError("shipped JMID checkpoint not found (set SNB_JMID_CHECKPOINT)")
    return synthetic_encoder(9), synthetic_ddpm(5), "seeded random-init weights of the JMID architecture (checkpoint not found)"
