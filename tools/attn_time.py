"""Times the attention kernel alone at the bench shape (512 envs x 4 heads x 1600 tokens), L2 flushed between launches.
SNB_ATTN_V2=1 selects the experimental second kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb import _capi  # noqa: E402

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
qkv = torch.randn(chunk, N, 1536, device="cuda").bfloat16()
out = torch.empty(chunk * N, 512, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
fn = lambda: _capi.check(_capi.lib.snb_jmid_attention(_capi.ptr(qkv), _capi.ptr(out), chunk, N, _capi.stream_ptr()), "attn")
for _ in range(3):
    fn()
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
t = float(np.mean(ts))
fl = 4.0 * N * N * 512 * chunk
print(f"attention {'v2' if os.environ.get('SNB_ATTN_V2') == '1' else 'v1'} {chunk} envs x {N} tokens: {t:.3f} ms, {fl / t / 1e9:.1f} TFLOP/s (min {fl / min(ts) / 1e9:.1f})")
