"""oracle/jmid_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU fp32 restatement (plain torch tensor algebra, no nn.Module) of the JMID / iMID
denoiser used by the reference's sim-inference path.  Each function cites the reference
file:line it follows (paths relative to the reference root).  Pinned against the
reference itself: tests/golden/jmid_*.npz were produced by oracle/gen_golden.py by running
sicnav_diffusion/JMID/MID/models/diffusion.py with (a) the shipped checkpoints and
(b) seeded synthetic weights from `make_random_weights` loaded into the reference module.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import math

import numpy as np
import torch

D_MODEL = 512
N_HEAD = 4
D_FF = 1024
CTX_DIM = 256
N_LAYER = 3


def variance_schedule(num_steps=100, beta_1=1e-4, beta_T=5e-2):
    """VarianceSchedule(mode='linear'), models/diffusion.py:12-56.  Returns fp32 tensors [101]."""
    betas = torch.linspace(beta_1, beta_T, steps=num_steps)
    betas = torch.cat([torch.zeros([1]), betas], dim=0)
    alphas = 1 - betas
    log_alphas = torch.log(alphas)
    for i in range(1, log_alphas.size(0)):
        log_alphas[i] += log_alphas[i - 1]
    alpha_bars = log_alphas.exp()
    return betas, alphas, alpha_bars


def weight_shapes():
    """Tensor names/shapes used by one forward of (Joint)TransformerConcatLinear(2, 256, 3)
    (models/diffusion.py:153-172, models/common.py:58-63)."""
    sh = {}

    def csl(name, din, dout):
        sh[f"net.{name}._layer.weight"] = (dout, din)
        sh[f"net.{name}._layer.bias"] = (dout,)
        sh[f"net.{name}._hyper_bias.weight"] = (dout, CTX_DIM + 3)
        sh[f"net.{name}._hyper_gate.weight"] = (dout, CTX_DIM + 3)
        sh[f"net.{name}._hyper_gate.bias"] = (dout,)

    csl("concat1", 2, D_MODEL)
    for l in range(N_LAYER):
        p = f"net.transformer_encoder.layers.{l}."
        sh[p + "self_attn.in_proj_weight"] = (3 * D_MODEL, D_MODEL)
        sh[p + "self_attn.in_proj_bias"] = (3 * D_MODEL,)
        sh[p + "self_attn.out_proj.weight"] = (D_MODEL, D_MODEL)
        sh[p + "self_attn.out_proj.bias"] = (D_MODEL,)
        sh[p + "linear1.weight"] = (D_FF, D_MODEL)
        sh[p + "linear1.bias"] = (D_FF,)
        sh[p + "linear2.weight"] = (D_MODEL, D_FF)
        sh[p + "linear2.bias"] = (D_MODEL,)
        for n in ("norm1", "norm2"):
            sh[p + n + ".weight"] = (D_MODEL,)
            sh[p + n + ".bias"] = (D_MODEL,)
    csl("concat3", D_MODEL, CTX_DIM)
    csl("concat4", CTX_DIM, CTX_DIM // 2)
    csl("linear", CTX_DIM // 2, 2)
    return sh


def make_random_weights(seed=0):
    """Seeded synthetic weights (numpy PCG64, reproducible on any box without the checkpoint).
    Uniform(+-1/sqrt(fan_in)) like nn.Linear's default; LayerNorm weight 1 +- 0.1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    for k, s in weight_shapes().items():
        if ".norm" in k:
            v = (1.0 + 0.1 * rng.standard_normal(s)) if k.endswith("weight") else 0.05 * rng.standard_normal(s)
        else:
            b = 1.0 / math.sqrt(s[1]) if len(s) == 2 else 0.05
            v = rng.uniform(-b, b, s)
        w[k] = torch.from_numpy(np.asarray(v, np.float32))
    return w


def positional_encoding(T, d_model=D_MODEL):
    """PositionalEncoding buffer rows 0..T-1, models/common.py:37-51."""
    pe = torch.zeros(T, d_model)
    position = torch.arange(0, T, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _csl(w, name, ctx_emb, x):
    """ConcatSquashLinear.forward, models/common.py:65-72.  ctx_emb [R,1,259], x [R,T,din]."""
    gate = torch.sigmoid(ctx_emb @ w[f"net.{name}._hyper_gate.weight"].T + w[f"net.{name}._hyper_gate.bias"])
    bias = ctx_emb @ w[f"net.{name}._hyper_bias.weight"].T
    return (x @ w[f"net.{name}._layer.weight"].T + w[f"net.{name}._layer.bias"]) * gate + bias


def _layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def _encoder_layer(w, l, h):
    """nn.TransformerEncoderLayer(d=512, nhead=4, ff=1024), post-norm, ReLU, eval mode
    (models/diffusion.py:161-166).  h [L, 512] = ONE sequence of L tokens."""
    p = f"net.transformer_encoder.layers.{l}."
    L = h.shape[0]
    hd = D_MODEL // N_HEAD
    qkv = h @ w[p + "self_attn.in_proj_weight"].T + w[p + "self_attn.in_proj_bias"]
    q, k, v = qkv[:, :D_MODEL], qkv[:, D_MODEL:2 * D_MODEL], qkv[:, 2 * D_MODEL:]
    q = q.reshape(L, N_HEAD, hd).transpose(0, 1)
    k = k.reshape(L, N_HEAD, hd).transpose(0, 1)
    v = v.reshape(L, N_HEAD, hd).transpose(0, 1)
    att = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(0, 1).reshape(L, D_MODEL)
    o = o @ w[p + "self_attn.out_proj.weight"].T + w[p + "self_attn.out_proj.bias"]
    y = _layer_norm(h + o, w[p + "norm1.weight"], w[p + "norm1.bias"])
    f = torch.relu(y @ w[p + "linear1.weight"].T + w[p + "linear1.bias"])
    f = f @ w[p + "linear2.weight"].T + w[p + "linear2.bias"]
    return _layer_norm(y + f, w[p + "norm2.weight"], w[p + "norm2.bias"])


def net_forward(w, x, beta, context, joint=True):
    """JointPredictionTransformerConcatLinear.forward (models/diffusion.py:173-209, mask=None) when
    joint=True, TransformerConcatLinear.forward (:133-150) when joint=False.
    x [R,T,2], beta [R], context [R,256] -> eps [R,T,2]."""
    R, T, _ = x.shape
    beta = beta.view(R, 1, 1)
    context = context.view(R, 1, -1)
    time_emb = torch.cat([beta, torch.sin(beta), torch.cos(beta)], dim=-1)
    ctx_emb = torch.cat([time_emb, context], dim=-1)              # [R,1,259]
    h = _csl(w, "concat1", ctx_emb, x)                            # [R,T,512]
    h = h.permute(1, 0, 2) + positional_encoding(T).unsqueeze(1)  # [T,R,512]
    if joint:
        seq = h.reshape(T * R, D_MODEL)                           # token index tau*R + r  (:197-199)
        for l in range(N_LAYER):
            seq = _encoder_layer(w, l, seq)
        trans = seq.reshape(T, R, D_MODEL).permute(1, 0, 2)
    else:
        outs = []
        for r in range(R):                                        # R independent sequences of T tokens (:147)
            seq = h[:, r, :]
            for l in range(N_LAYER):
                seq = _encoder_layer(w, l, seq)
            outs.append(seq)
        trans = torch.stack(outs, 0)
    trans = _csl(w, "concat3", ctx_emb, trans)
    trans = _csl(w, "concat4", ctx_emb, trans)
    return _csl(w, "linear", ctx_emb, trans)


def ddim_timesteps(step, num_steps=100):
    """`for t in range(num_steps, 0, -stride)` with stride=int(100/step), models/diffusion.py:507-508."""
    stride = int(100 / step)
    return list(range(num_steps, 0, -stride)), stride


def sample(w, context, x_T, step=20, joint=True, sampling="ddim"):
    """DiffusionTraj.sample_sicnav_inference (models/diffusion.py:478-541) with x_T injected
    (the reference draws it with torch.randn, quirk q2).  context [A,256], x_T [S*A,T,2] with
    row r = s*A + a (ctx.repeat(S,1), :496).  Returns velocities [S,A,T,2]."""
    betas, alphas, alpha_bars = variance_schedule()
    A = context.shape[0]
    R = x_T.shape[0]
    S = R // A
    ctx = context.repeat(S, 1)
    x_t = x_T.clone()
    ts, stride = ddim_timesteps(step)
    for t in ts:
        alpha_bar = alpha_bars[t]
        alpha_bar_next = alpha_bars[t - stride]
        beta = betas[[t] * R]
        e = net_forward(w, x_t, beta, ctx, joint=joint)
        if sampling == "ddim":
            x0_t = (x_t - e * (1 - alpha_bar).sqrt()) / alpha_bar.sqrt()
            x_t = alpha_bar_next.sqrt() * x0_t + (1 - alpha_bar_next).sqrt() * e
        else:
            raise NotImplementedError("shipped sim configs use ddim (test_time_configs/mid_jp.yaml:39)")
    return x_t.reshape(S, A, -1, 2)


def integrate(v, p0, dt=0.25):
    """SingleIntegrator.integrate_samples: cumsum(v, dim=2)*dt + p0 (single_integrator.py:290-321).
    v [S,A,T,2], p0 [A,2] -> positions [S,A,T,2]."""
    return torch.cumsum(v, dim=2) * dt + p0.view(1, -1, 1, 2)
