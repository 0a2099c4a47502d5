"""CPU: the pandas-free history join / resampling of the B = 1 predictor object (snb/jmid/history.py) against the reference's own
pandas calls (oracle/history_oracle.py restates mid_sim_wrapper.py:244-298 with pandas itself)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pd = pytest.importorskip("pandas")
import history_oracle as HO  # noqa: E402

_spec = importlib.util.spec_from_file_location("snb_history", os.path.join(ROOT, "safe-interactive-crowdnav_b200", "snb", "jmid", "history.py"))
HI = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(HI)


def _lists(rng, H, times, drop=0.0):
    hum = [[[float(rng.normal()), float(rng.normal()), float(t)] for t in times if rng.random() >= drop] for _ in range(H)]
    rob = [[float(rng.normal()), float(rng.normal()), float(t)] for t in times if rng.random() >= drop]
    return hum, rob


def _check(hum, rob, dt, F=6):
    a_h, a_r = HI.resample_histories(hum, rob, dt, F)
    b_h, b_r = HO.gen_agent_frames(hum, rob, dt, F)
    assert a_h.shape == b_h.shape and a_r.shape == b_r.shape, (a_h.shape, b_h.shape)
    assert np.allclose(a_h, b_h, rtol=0, atol=1e-12) and np.allclose(a_r, b_r, rtol=0, atol=1e-12)
    return a_h, a_r


def test_frames_one_time_step_apart_are_returned_unchanged():
    rng = np.random.default_rng(0)
    for H, n in ((1, 6), (3, 6), (10, 9), (5, 3)):
        times = -2.5 + 0.25 * np.arange(n)
        hum, rob = _lists(rng, H, times)
        h, r = _check(hum, rob, 0.25)
        want = np.asarray(hum)[:, -6:, :2]
        assert np.array_equal(h, want) and np.array_equal(r, np.asarray(rob)[-6:, :2])


@pytest.mark.parametrize("seed", range(12))
def test_faster_recording_missing_frames_and_gaps_match_pandas(seed):
    rng = np.random.default_rng(100 + seed)
    H = int(rng.integers(1, 7))
    dt = [0.25, 0.2, 0.4, 0.1][seed % 4]
    kind = seed % 3
    if kind == 0:                                   # 20 Hz recording, resampled down to time_step
        times = 3.0 + 0.05 * np.arange(int(rng.integers(20, 60)))
    elif kind == 1:                                 # jittered stamps with long gaps (empty windows -> interpolation)
        times = np.cumsum(rng.choice([0.03, 0.07, 0.25, 0.61, 1.3], int(rng.integers(8, 30))))
    else:                                           # one frame per time_step, negative start like the simulator's warm-up clock
        times = -2.5 + dt * np.arange(int(rng.integers(4, 12)))
    hum, rob = _lists(rng, H, times, drop=0.15 if seed % 2 else 0.0)
    if not rob or any(not h for h in hum):
        pytest.skip("an agent lost every frame")
    common = set(t for _, _, t in rob)
    for h in hum:
        common &= set(t for _, _, t in h)
    if not common:
        pytest.skip("no common time stamp")
    _check(hum, rob, dt)


def test_truncation_of_time_times_100_follows_pandas():
    # 0.1 * 3 * 100 = 30.000000000000004 -> 30 ns; 0.29 * 100 = 28.999999999999996 -> 28 ns: a different window than rounding would give
    times = [0.0, 0.1 * 3, 0.29, 0.55, 0.58, 0.8]
    rng = np.random.default_rng(5)
    hum, rob = _lists(rng, 2, times)
    _check(hum, rob, 0.25)


def test_random_recordings_match_pandas_hypothesis():
    """Property form: arbitrary (sorted, unique) time stamps on a 10 ms grid, arbitrary drops per agent."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=60, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(ticks=st.lists(st.integers(min_value=-400, max_value=2000), min_size=2, max_size=40, unique=True),
               H=st.integers(min_value=1, max_value=5), dt=st.sampled_from([0.25, 0.2, 0.1, 0.5]), seed=st.integers(0, 2 ** 16),
               drop=st.sampled_from([0.0, 0.1, 0.3]))
    def run(ticks, H, dt, seed, drop):
        rng = np.random.default_rng(seed)
        times = np.sort(np.asarray(ticks, np.float64)) * 0.01
        hum, rob = _lists(rng, H, times, drop=drop)
        if not rob or any(not h for h in hum):
            return
        common = set(t for _, _, t in rob)
        for h in hum:
            common &= set(t for _, _, t in h)
        if not common or min(t for _, _, t in hum[0]) not in [t for _, _, t in hum[0]]:
            return
        _check(hum, rob, dt)

    run()
