"""One ORCA crowd step at an HBM-sized batch (for ncu): python tools/crowd_prof.py <log2 envs> [steps]"""
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

B = 1 << int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = configparser.RawConfigParser()
cfg.read_string(ENV_CFG.format(H=10))
env = CrowdSimPlusBatch(B, "cuda")
env.LOG_DEPTH = 0
env.configure(cfg)
env.freeze_done = False
env.reset('test', test_cases=np.arange(B) % 500)
act = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); act[:, 1] = 0.5
flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
ts = []
for k in range(steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env._launch(act, None); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
t = float(np.mean(ts[2:]))
print(f"ORCA {B} x 10: {t * 1e3:.1f} us per step, {B / (t * 1e-3) / 1e6:.1f} M env-steps/s  (all: {[round(x, 3) for x in ts]})")
if len(sys.argv) > 3:      # a whole episode, every launch timed
    env.reset('test', test_cases=np.arange(B) % 500)
    ep = []
    for k in range(int(sys.argv[3])):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env._launch(act, None); b.record(); torch.cuda.synchronize()
        ep.append(a.elapsed_time(b))
    ep = np.asarray(ep)
    print(f"episode of {len(ep)} steps: mean {ep.mean() * 1e3:.1f} us = {B / (ep.mean() * 1e-3) / 1e6:.1f} M env-steps/s; min {ep.min():.3f} max {ep.max():.3f} ms at step {ep.argmax()}; "
          f"by decade {[round(float(ep[k:k + 10].mean()), 2) for k in range(0, len(ep), 10)]}")
