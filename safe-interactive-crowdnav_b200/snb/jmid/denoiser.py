"""JmidDenoiser -- Python handle of the sm_100a JMID / iMID denoiser (snb_jmid_* in include/snb.h).

Replaces, for the sim-inference path, `DiffusionTraj.sample_sicnav_inference` + the noise network
(sicnav_diffusion/JMID/MID/models/diffusion.py:153-209, 478-541).  torch tensors only hold the fp32 weights (the
reference state_dict layout) and the I/O buffers; all arithmetic is in libsnb.so.
"""
import ctypes as C
import math

import torch

from .. import _capi

CSL = ("concat1", "concat3", "concat4", "linear")


def positional_encoding(T, d_model=512):
    """rows 0..T-1 of PositionalEncoding.pe (models/common.py:37-51); used when the state_dict has no buffer."""
    pe = torch.zeros(T, d_model)
    position = torch.arange(0, T, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def variance_schedule(num_steps=100, beta_1=1e-4, beta_T=5e-2):
    """VarianceSchedule(mode='linear') buffers (models/diffusion.py:12-56); used when the state_dict has none."""
    betas = torch.cat([torch.zeros([1]), torch.linspace(beta_1, beta_T, steps=num_steps)], dim=0)
    log_alphas = torch.log(1 - betas)
    for i in range(1, log_alphas.size(0)):
        log_alphas[i] += log_alphas[i - 1]
    return betas, log_alphas.exp()


def weights_struct(sd, device, T):
    """state_dict (keys `net.*`, `var_sched.*`; a `vel_predictor.` prefix is stripped) -> (SnbJmidWeights, keepalive)."""
    sd = {(k[len("vel_predictor."):] if k.startswith("vel_predictor.") else k): v for k, v in sd.items()}
    keep = []

    def dev(t):
        t = t.detach().to(device=device, dtype=torch.float32).contiguous()
        keep.append(t)
        return t.data_ptr()

    w = _capi.JmidWeights()
    for name in CSL:
        c = getattr(w, name)
        c.layer_w = dev(sd[f"net.{name}._layer.weight"]); c.layer_b = dev(sd[f"net.{name}._layer.bias"])
        c.hyper_bias_w = dev(sd[f"net.{name}._hyper_bias.weight"])
        c.hyper_gate_w = dev(sd[f"net.{name}._hyper_gate.weight"]); c.hyper_gate_b = dev(sd[f"net.{name}._hyper_gate.bias"])
    for l in range(3):
        p = f"net.transformer_encoder.layers.{l}."
        e = w.layers[l]
        e.in_proj_w = dev(sd[p + "self_attn.in_proj_weight"]); e.in_proj_b = dev(sd[p + "self_attn.in_proj_bias"])
        e.out_proj_w = dev(sd[p + "self_attn.out_proj.weight"]); e.out_proj_b = dev(sd[p + "self_attn.out_proj.bias"])
        e.lin1_w = dev(sd[p + "linear1.weight"]); e.lin1_b = dev(sd[p + "linear1.bias"])
        e.lin2_w = dev(sd[p + "linear2.weight"]); e.lin2_b = dev(sd[p + "linear2.bias"])
        e.norm1_w = dev(sd[p + "norm1.weight"]); e.norm1_b = dev(sd[p + "norm1.bias"])
        e.norm2_w = dev(sd[p + "norm2.weight"]); e.norm2_b = dev(sd[p + "norm2.bias"])
    pe = sd["net.pos_emb.pe"].reshape(-1, 512)[:T] if "net.pos_emb.pe" in sd else positional_encoding(T)
    w.pos_emb = dev(pe)
    if "var_sched.betas" in sd:
        betas, abar = sd["var_sched.betas"], sd["var_sched.alpha_bars"]
    else:
        betas, abar = variance_schedule()
    w.betas = dev(betas); w.alpha_bars = dev(abar)
    return w, keep


class JmidDenoiser:
    PRECISIONS = {"bf16": 0, "fp32x": 1}

    def __init__(self, state_dict, max_envs, A, S, T=8, joint=True, device="cuda", precision="bf16"):
        if _capi.lib.snb_jmid_create is None:
            raise _capi.SnbError("libsnb.so was built without the denoiser")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.SnbError("JmidDenoiser needs a CUDA device (snb has no CPU path)")
        self.A, self.S, self.T, self.joint, self.max_envs = int(A), int(S), int(T), bool(joint), int(max_envs)
        w, keep = weights_struct(state_dict, self.device, self.T)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.snb_jmid_create(C.byref(self._h), C.byref(w), self.max_envs, self.A, self.S, self.T,
                                                  int(self.joint), _capi.stream_ptr()), "snb_jmid_create")
            torch.cuda.current_stream().synchronize()
        del keep  # the library copied / converted everything it needs
        self.precision = "bf16"
        if precision != "bf16":
            self.set_precision(precision)

    def set_precision(self, precision):
        """"bf16" (default: bf16 tensor-core operands, fp32 accumulation) or "fp32x" (fp32-class: split-bf16 GEMMs + fp32 SIMT attention,
        the parity instrument against the reference's fp32 path; ~6x the FLOPs, 8 environments per chunk)."""
        if precision not in self.PRECISIONS:
            raise _capi.SnbError(f"unknown precision {precision!r} (bf16 | fp32x)")
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.snb_jmid_set_precision(self._h, self.PRECISIONS[precision], _capi.stream_ptr()), "snb_jmid_set_precision")
        self.precision = precision

    def _chk(self, t, shape):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == tuple(shape), (t.shape, shape)

    def denoise(self, ctx, x_T, n_steps=20, out=None, stream=None):
        """ctx [B,A',256], x_T [B,S*A',T,2] (row r = s*A' + a) -> velocities [B,S,A',T,2] (all fp32 CUDA); A' <= A, the
        number of agents of this call (the attention cluster of the predictor changes size from step to step)."""
        B, A = ctx.shape[0], ctx.shape[1]
        if not 1 <= A <= self.A:
            raise _capi.SnbError(f"denoise: {A} agents per environment, the handle was built for at most {self.A}")
        self._chk(ctx, (B, A, 256)); self._chk(x_T, (B, self.S * A, self.T, 2))
        if out is None:
            out = torch.empty(B, self.S, A, self.T, 2, dtype=torch.float32, device=self.device)
        _capi.check(_capi.lib.snb_jmid_denoise_agents(self._h, _capi.ptr(ctx), _capi.ptr(x_T), _capi.ptr(out), B, A, int(n_steps),
                                                      _capi.stream_ptr(stream)), "snb_jmid_denoise")
        return out

    def eps(self, ctx, x_t, t, stream=None):
        """One noise-network forward at diffusion step t: returns eps [B,S*A,T,2]."""
        B = ctx.shape[0]
        self._chk(ctx, (B, self.A, 256)); self._chk(x_t, (B, self.S * self.A, self.T, 2))
        out = torch.empty_like(x_t)
        _capi.check(_capi.lib.snb_jmid_eps(self._h, _capi.ptr(ctx), _capi.ptr(x_t), _capi.ptr(out), B, int(t),
                                           _capi.stream_ptr(stream)), "snb_jmid_eps")
        return out

    def integrate(self, vel, p0, dt=0.25, stream=None):
        """vel [B,S,A,T,2], p0 [B,A,2] -> positions [B,S,A,T,2] (SingleIntegrator.integrate_samples)."""
        B, A = vel.shape[0], vel.shape[2]
        self._chk(vel, (B, self.S, A, self.T, 2)); self._chk(p0, (B, A, 2))
        pos = torch.empty_like(vel)
        _capi.check(_capi.lib.snb_jmid_integrate(_capi.ptr(vel), _capi.ptr(p0), _capi.ptr(pos), B, self.S, A, self.T,
                                                 float(dt), _capi.stream_ptr(stream)), "snb_jmid_integrate")
        return pos

    def predict_host(self, ctx_np, x_T_np, p0_np, n_steps=20, dt=0.25):
        """Host buffers in, host positions out (H2D -> denoise -> integrate -> D2H inside the call)."""
        import numpy as np
        ctx_np = np.ascontiguousarray(ctx_np, np.float32); x_T_np = np.ascontiguousarray(x_T_np, np.float32)
        p0_np = np.ascontiguousarray(p0_np, np.float32)
        B = ctx_np.shape[0]
        pos = np.empty((B, self.S, self.A, self.T, 2), np.float32)
        fp = C.POINTER(C.c_float)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.snb_jmid_predict_host(self._h, ctx_np.ctypes.data_as(fp), x_T_np.ctypes.data_as(fp),
                                                        p0_np.ctypes.data_as(fp), pos.ctypes.data_as(fp), B, int(n_steps),
                                                        float(dt)), "snb_jmid_predict_host")
        return pos

    def flops_per_iter(self):
        return float(_capi.lib.snb_jmid_flops_per_iter(self.A, self.S, self.T, int(self.joint)))

    def __del__(self):
        h = getattr(self, "_h", None)
        lib = getattr(_capi, "lib", None)      # None while the interpreter shuts down
        if h and lib is not None:
            lib.snb_jmid_destroy(h)
            self._h = None
