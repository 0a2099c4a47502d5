"""GPU: episode rollouts (snb/rollout.py, BASELINE configs[4]) against the CPU restatement oracle/rollout_oracle.py (C step oracle +
scenario oracle): ORCA arithmetic is bit-exact between the kernel and the oracle, so with the same robot actions every step's flag
word and the episode counters simple_test.py pickles must be IDENTICAL, for 256 test cases run to done / time limit."""
import configparser

import numpy as np
import pytest

import rollout_oracle as RO

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CFG = """
[env]
time_limit = {time_limit}
time_step = 0.25
val_size = 100
test_size = 500
randomize_attributes = true
[sim]
train_val_sim = circle_crossing
test_sim = circle_crossing
starts_moving = 10
square_width = 5
circle_radius = 4.0
rect_width = 1.75
rect_height = 4
human_num = {H}
[humans]
visible = true
policy = orca
radius = 0.3
sensor = coordinates
safety_space = 0.05
v_pref = 1.5
[robot]
visible = true
policy = linear
radius = 0.25
v_pref = 1.0
sensor = coordinates
[reward]
success_reward = 1
collision_penalty = -0.25
freezing_penalty = -0.125
discomfort_dist = 0.2
discomfort_penalty_factor = 0.5
"""


def _env(B, H, time_limit=30):
    from snb.env import CrowdSimPlusBatch
    cfg = configparser.RawConfigParser()
    cfg.read_string(CFG.format(H=H, time_limit=time_limit))
    env = CrowdSimPlusBatch(B, "cuda")
    env.configure(cfg)
    return env


@pytest.mark.parametrize("H,time_limit", [(10, 30), (5, 6)])
def test_orca_episode_metrics_identical_to_the_oracle(H, time_limit):
    from snb import rollout
    B = 256
    cases = list(range(B))
    env = _env(B, H, time_limit)
    # ORCA.configure(config) takes one argument, so the human policy keeps safety_space = 0 (quirk q6, human_plus.py:12-16)
    oenv, pcfg, dcfg, rcfg = RO.reset(cases, H, time_limit=float(time_limit), reward=dict(collision_penalty=-0.25, freezing_penalty=-0.125))
    acts, gflags, gdmin, gactive = [], [], [], []

    def on_step(k, e):
        acts.append(robot.action.cpu().numpy().copy())
        gflags.append(e.flags.cpu().numpy().copy()); gdmin.append(e.dmin.cpu().numpy().copy())

    robot = rollout.LinearRobot(env)
    em, steps, env_steps, _ = rollout.run_episodes(env, cases, robot=robot, on_step=on_step, poll_every=1)
    # the device-built scenes after the warm-up equal the oracle's (positions to 1e-9: libm ulps in the generator)
    m_ref, trace = RO.run_episodes(oenv, pcfg, dcfg, rcfg, actions=lambda k, e: acts[k])
    assert steps == len(trace)
    for k, (f, d, live) in enumerate(trace):
        assert np.array_equal(gflags[k][live], f[live]), k
        assert np.array_equal(gflags[k][~live], np.zeros((~live).sum(), np.int32)), k       # frozen environments report nothing
        fin = np.isfinite(d) & live
        assert np.allclose(gdmin[k][fin], d[fin], atol=1e-9), k
    m = em.m.cpu().numpy()
    assert np.array_equal(m[:, :8], m_ref[:, :8])
    assert np.allclose(m[:, 8], m_ref[:, 8], atol=1e-9)
    assert env_steps == int(m_ref[:, 2].sum()) and int(em.live.sum().item()) == 0
    assert np.all((m[:, 0] + m[:, 1]) == 1.0)                       # every episode ended by reaching the goal or by the time limit
    s = rollout.summarize(em.m)
    assert s["episodes"] == B and abs(s["success_rate"] + s["timeout_rate"] - 1.0) < 1e-12
    if time_limit == 6:
        assert s["timeout_rate"] > 0.5                              # 8 m to go at 1 m/s: the limit really is exercised
    # the Linear robot on the device is the reference's formula (libm ulps only)
    assert np.max(np.abs(acts[0] - np.array([[0.0, 1.0]]))) < 1e-12


def test_rollout_with_the_predictor_in_the_loop():
    """update_state_hists + predict_ret_best + MPC ingest before every robot action (sicnav_acados.py:1640-1667), history bootstrapped
    from the environment's state log at reset (:1163-1182): shapes, finiteness, and the forecasts' current pose == the humans' positions."""
    from snb import rollout
    from snb.jmid.forecaster import ForecasterBatch
    from snb.jmid.weights import default_weights
    B, H, S = 8, 5, 6
    env = _env(B, H, 3)
    enc, ddpm, _ = default_weights()
    fc_ = ForecasterBatch(enc, ddpm, max_envs=B, H=H, num_samples=S, step_size=2)
    seen = []
    em, steps, env_steps, ingest = rollout.run_episodes(env, list(range(B)), forecaster=fc_, on_step=lambda k, e: seen.append(k))
    resh, wts, goals, vpref = ingest
    assert steps == len(seen) and steps >= 12 and tuple(resh.shape) == (B, 5, H * S, 2) and tuple(wts.shape) == (B, S)
    assert bool(torch.isfinite(resh).all()) and bool(torch.isfinite(goals).all()) and bool((vpref >= 0).all())
    s = rollout.summarize(em.m)
    assert s["timeout_rate"] == 1.0 and s["mean_steps"] == 13.0      # time_limit 3 s: done on the step that starts at t = 3.0
