"""The PCG64 / SeedSequence restatement behind the device scenario reset (oracle/pcg64_oracle.py, same arithmetic as
csrc/scene_kernels.cu) against numpy itself -- the generator the reference seeds every episode with
(crowd_sim_plus.py:658-664) -- and as a drop-in `rng` for the host scene generator."""
import numpy as np
import pytest

import pcg64_oracle as po


@pytest.mark.parametrize("seed", [0, 1, 999, 1000, 1001, 1499, 2000, 123456789, 2 ** 32 - 1, 2 ** 32 + 5, 2 ** 63 + 11])
def test_state_and_stream_equal_numpy(seed):
    st = np.random.PCG64(seed).state["state"]
    g = po.Pcg64(seed)
    assert g.state == st["state"] and g.inc == st["inc"]
    rng = np.random.default_rng(seed)
    for _ in range(64):
        assert g.random() == rng.random()
    for _ in range(16):
        assert g.uniform(0.5, 1.5) == rng.uniform(0.5, 1.5)


@pytest.mark.parametrize("rule,H", [("circle_crossing", 10), ("hallway", 6), ("hallway_static", 5), ("hallway_bottleneck", 5)])
def test_scene_generator_on_the_restated_stream_equals_numpy_stream(rule, H):
    import scenario_oracle as scenario
    p = scenario.SceneParams(4.0, 2.5, 4, 0.3, 1.5, 0.25, 0.2, True)
    segs, door = scenario.static_obstacles(rule, p)
    for case in (0, 7, 123):
        a = scenario._generate_with_goals(rule, H, np.random.default_rng(1000 + case), p, segs, door)
        b = scenario._generate_with_goals(rule, H, po.Pcg64(1000 + case), p, segs, door)
        assert a == b
