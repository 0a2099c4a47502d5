"""Bring-up diagnostics for the tensor-core kernels (run on the GPU box): prints error statistics per component
instead of asserting, so one gpurun call shows the whole picture."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from snb import _capi  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def gemm(M, N, K, epi):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    resid = torch.randn(M, N, device=dev).bfloat16()
    out = torch.full((M, N), 7.0, device=dev, dtype=torch.float32 if epi == 2 else torch.bfloat16)
    rc = _capi.lib.snb_jmid_gemm_bf16(_capi.ptr(A), _capi.ptr(W), _capi.ptr(bias), _capi.ptr(out), M, N, K, epi,
                                      _capi.stream_ptr())
    torch.cuda.synchronize()
    ref = A.float() @ W.float().T + bias
    if epi == 1:
        ref = torch.relu(ref)
    err = (out.float() - ref).abs()
    print(f"gemm M={M} N={N} K={K} epi={epi}: rc={rc} max_err={err.max().item():.4e} mean_err={err.mean().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3f}")
    if err.max().item() > 0.1:
        bad = (err > 0.1)
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        print("   bad rows", rows, "bad cols", cols, "frac bad", bad.float().mean().item())
        print("   out[0,:8]", out[0, :8].float().tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())


def attention(n_env, n_tok):
    qkv = (torch.randn(n_env, n_tok, 1536, device=dev) * 1.0).bfloat16()
    out = torch.zeros(n_env * n_tok, 512, device=dev, dtype=torch.bfloat16)
    rc = _capi.lib.snb_jmid_attention(_capi.ptr(qkv), _capi.ptr(out), n_env, n_tok, _capi.stream_ptr())
    torch.cuda.synchronize()
    q, k, v = qkv.float().split(512, dim=-1)
    q = q.view(n_env, n_tok, 4, 128).transpose(1, 2); k = k.view(n_env, n_tok, 4, 128).transpose(1, 2)
    v = v.view(n_env, n_tok, 4, 128).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / 128 ** 0.5, -1)
    ref = (att @ v).transpose(1, 2).reshape(n_env * n_tok, 512)
    err = (out.float() - ref).abs()
    print(f"attn n_env={n_env} n_tok={n_tok}: rc={rc} max_err={err.max().item():.4e} mean_err={err.mean().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3f} nan={torch.isnan(out.float()).sum().item()}")
    if err.max().item() > 0.05:
        e2 = err.view(n_env, n_tok, 4, 128)
        print("   per-head max", e2.amax(dim=(0, 1, 3)).tolist())
        print("   per-env max", e2.amax(dim=(1, 2, 3)).tolist())
        print("   per-row-block(128) max", [e2[:, i:i + 128].max().item() for i in range(0, n_tok, 128)][:14])
        print("   per-col-block(32) max", [e2[..., i:i + 32].max().item() for i in range(0, 128, 32)])
        print("   out[0,:6]", out[0, :6].float().tolist(), "\n   ref[0,:6]", ref[0, :6].tolist())


def full(A, S, joint, steps, B=2, seed=5):
    import jmid_oracle as JO
    from snb.jmid import JmidDenoiser
    w = JO.make_random_weights(seed)
    g = torch.Generator().manual_seed(1)
    ctx = torch.randn(B, A, 256, generator=g)
    xT = torch.randn(B, S * A, 8, 2, generator=g)
    den = JmidDenoiser(w, max_envs=B, A=A, S=S, T=8, joint=joint)
    betas, _, _ = JO.variance_schedule()
    e = den.eps(ctx.cuda(), xT.cuda(), 55).cpu()
    with torch.no_grad():
        e_ref = torch.stack([JO.net_forward(w, xT[b], betas[[55] * (A * S)], ctx[b].repeat(S, 1), joint=joint) for b in range(B)])
    print(f"eps A={A} S={S} joint={joint}: max_err={(e - e_ref).abs().max().item():.4e} ref_absmax={e_ref.abs().max().item():.3f}")
    t0 = time.time()
    out = den.denoise(ctx.cuda(), xT.cuda(), n_steps=steps).cpu()
    with torch.no_grad():
        ref = torch.stack([JO.sample(w, ctx[b], xT[b], step=steps, joint=joint) for b in range(B)])
    print(f"denoise A={A} S={S} joint={joint} steps={steps}: max_err={(out - ref).abs().max().item():.4e} "
          f"ref_absmax={ref.abs().max().item():.3f} ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "attn", "full"]
    if "gemm" in what:
        for (M, N, K, epi) in ((128, 256, 64, 0), (256, 256, 128, 0), (391, 1536, 512, 0), (1600, 512, 1024, 2), (3200, 1024, 512, 1),
                               (1600, 128, 256, 0), (25600, 512, 512, 2)):
            gemm(M, N, K, epi)
    if "attn" in what:
        for (ne, nt) in ((1, 128), (2, 256), (2, 96), (1, 200), (2, 1600)):
            attention(ne, nt)
    if "full" in what:
        full(3, 4, True, 4)
        full(3, 4, False, 4)
        full(10, 20, True, 2, B=1)
    print("launches", _capi.launch_count())
