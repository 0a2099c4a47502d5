"""Debug: clock64() timeline of the second work item of CTA 0 of attn2_fwd_kernel.  Needs a library built with
SNB_NVCC_FLAGS=-DSNB_ATTN_TRACE python safe-interactive-crowdnav_b200/build.py --force"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb import _capi  # noqa: E402

chunk, N = int(os.environ.get("SNB_JMID_CHUNK", 512)), 1600
qkv = torch.randn(chunk, N, 1536, device="cuda").bfloat16()
out = torch.empty(chunk * N, 512, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _capi.check(_capi.lib.snb_jmid_attention(_capi.ptr(qkv), _capi.ptr(out), chunk, N, _capi.stream_ptr()), "attn")
torch.cuda.synchronize()
buf = np.zeros(3 * 32 * 8 + 8, dtype=np.int64)
assert _capi.lib.snb_debug_attn2_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
tr = buf[:3 * 32 * 8].reshape(3, 32, 8)
t0 = tr[tr > 0].min()
print("softmax: 0 wait-start 1 s_full 2 S loaded+released 3 exps done 4 pv(G-2) seen 5 P stored+arrived | j=31: 0 epilogue start 1 last PV seen 2 O written")
print("MMA: 0 K(j+1) ready 1 s_free[A] 2 S_A,S_B issued 3 V(j) ready 4 p_ready[A] 5 PV_A issued 6 p_ready[B] 7 PV_B issued")
for j in list(range(25)) + [31]:
    for r, nm in ((0, "A"), (1, "B"), (2, "MMA")):
        ev = tr[r, j]
        if (ev > 0).any():
            print(f"j={j:2d} {nm:4s} " + " ".join(f"{(int(e) - t0) if e > 0 else -1:7d}" for e in ev))
