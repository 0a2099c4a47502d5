"""One forward of the noise network at the bench chunk shape (512 envs x 1600 tokens) for ncu: every launch of one denoise
iteration once.  python tools/net_once.py [envs] [forwards]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
from snb.jmid import JmidDenoiser  # noqa: E402
from snb.jmid.weights import synthetic_ddpm  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
den = JmidDenoiser(synthetic_ddpm(5), max_envs=B, A=10, S=20, T=8, joint=True)
ctx = torch.randn(B, 10, 256, device="cuda"); x = torch.randn(B, 200, 8, 2, device="cuda")
for _ in range(n):
    den.eps(ctx, x, 55)
torch.cuda.synchronize()
print("done")
