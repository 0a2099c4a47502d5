"""Times / profiles the crowd step alone: ORCA 1024 x 10 (configs[1]) and an HBM-sized batch (ncu target)."""
import configparser
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-interactive-crowdnav_b200"))
sys.path.insert(0, ROOT)
from bench import ENV_CFG  # noqa: E402
from snb.env import CrowdSimPlusBatch  # noqa: E402

for B in (1024, 1 << 18):
    cfg = configparser.RawConfigParser()
    cfg.read_string(ENV_CFG.format(H=10))
    env = CrowdSimPlusBatch(B, "cuda")
    env.configure(cfg)
    env.freeze_done = False
    env.reset('test', test_cases=np.arange(B) % 500)
    act = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); act[:, 1] = 0.5
    for _ in range(3):
        env.step(act)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        env.step(act)
    b.record(); torch.cuda.synchronize()
    print(f"ORCA {B} x 10: {a.elapsed_time(b) / 10 * 1e3:.1f} us per step, {B / (a.elapsed_time(b) / 10 * 1e-3) / 1e6:.1f} M env-steps/s")
