"""Debug: per-block clock64() timeline of the attention kernel's CTA (0,0,0).  Needs a library built with
SNB_NVCC_FLAGS=-DSNB_ATTN_TRACE python safe-interactive-crowdnav_b200/build.py --force"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb import _capi  # noqa: E402

chunk, N = int(os.environ.get("SNB_JMID_CHUNK", 128)), 1600
qkv = torch.randn(chunk, N, 1536, device="cuda").bfloat16()
out = torch.empty(chunk * N, 512, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _capi.check(_capi.lib.snb_jmid_attention(_capi.ptr(qkv), _capi.ptr(out), chunk, N, _capi.stream_ptr()), "attn")
torch.cuda.synchronize()
buf = np.zeros(6 * 16 * 8, dtype=np.int64)
assert _capi.lib.snb_debug_attn_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
tr = buf.reshape(6, 16, 8)
cta = tr[5, 0]
tr = tr[:5]
t0 = cta[0]
print(f"CTA: setup done +{cta[1]-t0}, end +{cta[2]-t0} clk; {cta[4]-cta[3]} ns -> {(cta[2]-t0)/(cta[4]-cta[3])*1e3:.0f} MHz")
names = ["A", "B", "-", "-", "MMA"]
print("softmax: 0 wait-start 1 s_full 2 S loaded+released 3 pv_done(j-1) seen 4 exps+P stored 5 arrived;  MMA: 0 K(j+1)+s_free[A] 1 S_A issued 2 S_B issued 3 p_ready[A] 4 PV_A issued 5 p_ready[B] 6 PV_B issued")
for j in range(13):
    for r in (0, 1, 4):
        ev = tr[r, j]
        print(f"j={j:2d} {names[r]:5s} " + " ".join(f"{(int(e) - t0) if e > 0 else -1:7d}" for e in ev[:8]))
