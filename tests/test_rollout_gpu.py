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


def test_frozen_environments_report_zeros_and_keep_their_state():
    """freeze_done: an environment that finished is skipped by the kernel; its reward / flags read 0 on every later step (summing
    rewards over an episode must not re-count the terminal reward) and its state no longer moves."""
    B = 64
    env = _env(B, 5, 2)                    # time limit 2 s: every env times out on the step that starts at t = 2.0
    env.freeze_done = True
    env.reset('test', test_cases=list(range(B)))
    act = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); act[:, 1] = 0.3
    total = torch.zeros(B, dtype=torch.float64, device="cuda")
    total_live = torch.zeros(B, dtype=torch.float64, device="cuda")   # rewards of the steps up to and including the terminal one
    done_at = torch.full((B,), -1, dtype=torch.int64, device="cuda")
    for k in range(14):
        reward, done, flags = env.step(act)
        total += reward
        total_live += reward * (done_at < 0)
        done_at = torch.where((done_at < 0) & done, torch.full_like(done_at, k), done_at)
        if k == 10:
            px10 = env.state.px.clone()
    assert bool((done_at == 8).all())                                # t = 0, .25, ..., 2.0 -> the 9th step reports the time-out
    assert int(env.active.sum().item()) == 0
    assert bool((env.flags == 0).all()) and bool((env.reward == 0).all())      # later steps: nothing reported
    assert torch.equal(px10, env.state.px)                           # and nothing moved after the freeze
    assert torch.equal(total, total_live) and bool((total <= -1.0 + 1e-9).all())   # the terminal reward (time-out, -1) is counted once


def test_configure_refuses_what_it_does_not_implement():
    """the SB3 reward branch / smoothness penalties / occlusion are refused instead of silently returning other rewards;
    plain ORCA humans in a hallway raise like generate_hallway_human (crowd_sim_plus.py:524-525); is_bottleneck is per episode."""
    from snb.env import CrowdSimPlusBatch

    def cfg_with(**over):
        cfg = configparser.RawConfigParser()
        cfg.read_string(CFG.format(H=4, time_limit=5))
        for k, v in over.items():
            sec, key = k.split("__")
            cfg.set(sec, key, str(v))
        return cfg
    for bad in (dict(env__SB3="true"), dict(reward__angular_smoothness_factor="0.1"), dict(env__occlusion="true")):
        with pytest.raises(NotImplementedError):
            CrowdSimPlusBatch(4, "cuda").configure(cfg_with(**bad))
    env = CrowdSimPlusBatch(4, "cuda")
    env.configure(cfg_with(sim__test_sim="hallway", sim__train_val_sim="hallway"))
    with pytest.raises(RuntimeError, match="orca_plus or sfm"):
        env.reset('test', test_cases=[0, 1, 2, 3])
    sfm = dict(humans__policy="sfm", humans__A="3.0", humans__B="0.18", humans__KI="1.0", humans__A_static="2.0", humans__B_static="0.025",
               humans__A_bottleneck="6.0", humans__B_bottleneck="0.12", humans__radius="0.2", sim__rect_width="2.0", sim__circle_radius="1.5")
    env = CrowdSimPlusBatch(4, "cuda")
    env.configure(cfg_with(sim__test_sim="hallway_bottleneck", sim__train_val_sim="hallway", **sfm))
    env.reset('test', test_cases=[0, 1, 2, 3])
    assert env.human_policy.is_bottleneck is True
    env.reset('val', test_cases=[0, 1, 2, 3])                        # another layout with the SAME policy object
    assert env.human_policy.is_bottleneck is False


def test_orca_hbm_sized_batch_replicas_stay_bit_identical_and_match_the_oracle():
    """BASELINE-scale property (size independent): 2^18 environments cycle through 512 distinct test cases, so environment b and
    b mod 512 start from the same scene but sit on different lanes, warps and CTAs of the warp-owns-environments kernel.  After 24 steps
    through the congested middle of the crossing (where a third of the humans need linearProgram3) every replica must hold the SAME bits,
    and the first 512 must hold the bits of the CPU oracle stepping those scenes from the same post-reset state."""
    import oracle_lib as ol
    from snb import rollout
    B, H, R, steps = 1 << 18, 10, 512, 24
    env = _env(B, H, 30)
    env.freeze_done = False
    env.reset('test', test_cases=np.arange(B) % R)
    robot = rollout.LinearRobot(env)
    st = env.state
    for n in ("px", "py", "vx", "vy", "gx", "gy", "vpref"):
        a = getattr(st, n).view(B // R, R, H)
        assert torch.equal(a, a[0:1].expand_as(a)), n        # the device reset is replica-invariant too
    # the oracle starts from the device's post-reset state of the first R environments (the generators agree to libm ulps only)
    oenv, pcfg, dcfg, rcfg = RO.reset(list(range(R)), H, time_limit=30.0, starts_moving=0, reward=dict(collision_penalty=-0.25, freezing_penalty=-0.125))
    for n in ("px", "py", "vx", "vy", "theta", "gx", "gy", "fgx", "fgy", "vpref", "radius", "human_time"):
        getattr(oenv, n)[:] = getattr(st, n)[:R].cpu().numpy().ravel()
    for n, m in (("rpx", "ex_px"), ("rpy", "ex_py"), ("rvx", "ex_vx"), ("rvy", "ex_vy")):
        getattr(oenv, n)[:] = getattr(st, m)[:R].cpu().numpy().ravel()
    for n in ("rtheta", "rgx", "rgy", "global_time", "prev_dist"):
        getattr(oenv, n)[:] = getattr(st, n)[:R].cpu().numpy().ravel()
    for k in range(steps):
        act = robot.act()
        a_host = act[:R].cpu().numpy().copy()
        env.step(act)
        ol.env_step(pcfg, dcfg, rcfg, oenv, a_host, n_threads=8)
    torch.cuda.synchronize()
    env.check_status()
    for n in ("px", "py", "vx", "vy", "theta", "human_time"):
        a = getattr(st, n).view(B // R, R, H)
        assert torch.equal(a, a[0:1].expand_as(a)), n
    fl = env.flags.view(B // R, R)
    assert torch.equal(fl, fl[0:1].expand_as(fl))
    for n in ("px", "py", "vx", "vy"):
        got = getattr(st, n)[:R].cpu().numpy().ravel()
        assert np.array_equal(got, getattr(oenv, n)), (n, np.max(np.abs(got - getattr(oenv, n))))


@pytest.mark.parametrize("B", [64, 4096])
def test_state_log_written_by_the_step_launch(B):
    """CrowdSimPlus.step appends the state it starts from to self.states (crowd_sim_plus.py:1175-1181); here the ring slot is written by
    the step's own launch (snb_env_step_logged) -- by crowd_step_kernel for 64 environments, by crowd_orca_warp_kernel for 4096."""
    H = 10
    env = _env(B, H, 30)
    env.freeze_done = True                      # frozen environments are logged too
    env.reset('test', test_cases=np.arange(B) % 500)
    assert env.n_logged == 10                   # the starts_moving warm-up steps
    st = env.state
    act = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); act[:, 1] = 1.0
    env.active[::3] = 0
    for k in range(9):                          # wraps the 7-slot ring
        before = torch.cat([torch.stack([st.px, st.py], -1), torch.stack([st.ex_px, st.ex_py], -1)], 1).clone()   # [B, H + 1, 2]
        slot = env.n_logged % env.LOG_DEPTH
        env.step(act)
        assert torch.equal(env.state_log[:, slot], before), k
    env.check_status()
