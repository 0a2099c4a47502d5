import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb import _capi
M = 25600
A = torch.randn(M, 512, device="cuda").bfloat16(); W = (torch.randn(1536, 512, device="cuda") * 0.05).bfloat16(); b = torch.zeros(1536, device="cuda")
o = torch.empty(M, 1536, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    _capi.check(_capi.lib.snb_jmid_gemm_bf16(_capi.ptr(A), _capi.ptr(W), _capi.ptr(b), _capi.ptr(o), M, 1536, 512, 0, _capi.stream_ptr()), "gemm")
torch.cuda.synchronize()
