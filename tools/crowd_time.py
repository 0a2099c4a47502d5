"""Times the crowd step alone in both ORCA phase-1 modes (SNB_CROWD_MODE=warp|thread): configs[1] and larger batches."""
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

for B, H, pol, sim in ((1024, 10, "orca", "circle_crossing"), (4096, 10, "orca", "circle_crossing"), (1 << 18, 10, "orca", "circle_crossing"),
                       (1024, 25, "sfm", "hallway"), (4096, 25, "sfm", "hallway"), (1 << 16, 25, "sfm", "hallway")):
    cfg = configparser.RawConfigParser()
    txt = ENV_CFG.format(H=H).replace("policy = orca", f"policy = {pol}").replace("circle_crossing", sim)
    if sim == "hallway":
        txt = txt.replace("rect_width = 1.75", "rect_width = 6.0").replace("rect_height = 4", "rect_height = 12")
    cfg.read_string(txt)
    env = CrowdSimPlusBatch(B, "cuda")
    env.configure(cfg)
    env.freeze_done = False
    env.reset('test', test_cases=np.arange(B) % 500)
    act = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); act[:, 1] = 0.5
    flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
    for mode in ("warp", "thread"):
        os.environ["SNB_CROWD_MODE"] = mode
        for _ in range(3):
            env.step(act)
        ts = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.step(act); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.mean(ts))
        print(f"{pol:5s} {B:7d} x {H:2d}  {mode:6s}: {t * 1e3:9.1f} us per step, {B / (t * 1e-3) / 1e6:7.1f} M env-steps/s")
