"""Times the attention kernel and the GEMM shapes of one bench chunk (16 envs x 1600 tokens) with CUDA events."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb import _capi  # noqa: E402

dev = "cuda"
chunk, N = int(os.environ.get("SNB_JMID_CHUNK", 16)), 1600
M = chunk * N
flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)


def timeit(fn, reps=10, cold=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


qkv = torch.randn(chunk, N, 1536, device=dev).bfloat16()
out = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
for cold in (True, False):
    t, tmin = timeit(lambda: _capi.check(_capi.lib.snb_jmid_attention(_capi.ptr(qkv), _capi.ptr(out), chunk, N, _capi.stream_ptr()), "attn"), cold=cold)
    fl = 4.0 * N * N * 512 * chunk
    print(f"attention cold={cold}: {t * 1e3:.1f} us (min {tmin * 1e3:.1f})  {fl / t / 1e9:.0f} TFLOP/s")
for (n, k, epi) in [(1536, 512, 0), (512, 512, 0), (1024, 512, 1), (512, 1024, 0), (256, 512, 0), (128, 256, 0)]:
    A = torch.randn(M, k, device=dev).bfloat16(); W = (torch.randn(n, k, device=dev) * 0.05).bfloat16(); b = torch.zeros(n, device=dev)
    o = torch.empty(M, n, device=dev, dtype=torch.float32 if epi == 2 else torch.bfloat16)
    for cold in (True, False):
        t, tmin = timeit(lambda: _capi.check(_capi.lib.snb_jmid_gemm_bf16(_capi.ptr(A), _capi.ptr(W), _capi.ptr(b), _capi.ptr(o), M, n, k, epi,
                                                                         _capi.stream_ptr()), "gemm"), cold=cold)
        print(f"gemm N={n} K={k} epi={epi} cold={cold}: {t * 1e3:.1f} us (min {tmin * 1e3:.1f})  {2.0 * M * n * k / t / 1e9:.0f} TFLOP/s")
