"""Short driver for ncu: runs the two tensor-core kernels at the shapes of one bench chunk (16 envs x 1600 tokens)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb import _capi  # noqa: E402

dev = "cuda"
chunk, N = int(os.environ.get("SNB_JMID_CHUNK", 16)), 1600
M = chunk * N
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
qkv = torch.randn(chunk, N, 1536, device=dev).bfloat16()
out = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
shapes = [(1536, 512, 0), (512, 512, 2), (1024, 512, 1), (512, 1024, 2), (256, 512, 0), (128, 256, 0)]
bufs = []
for (n, k, epi) in shapes:
    bufs.append((torch.randn(M, k, device=dev).bfloat16(), torch.randn(n, k, device=dev).bfloat16() * 0.05, torch.zeros(n, device=dev),
                 torch.randn(M, n, device=dev).bfloat16(), torch.empty(M, n, device=dev, dtype=torch.float32 if epi == 2 else torch.bfloat16)))
for _ in range(reps):
    _capi.check(_capi.lib.snb_jmid_attention(_capi.ptr(qkv), _capi.ptr(out), chunk, N, _capi.stream_ptr()), "attn")
    for (n, k, epi), (A, W, b, r, o) in zip(shapes, bufs):
        _capi.check(_capi.lib.snb_jmid_gemm_bf16(_capi.ptr(A), _capi.ptr(W), _capi.ptr(b), _capi.ptr(o), M, n, k, epi,
                                                 _capi.stream_ptr()), "gemm")
torch.cuda.synchronize()
print("done")
