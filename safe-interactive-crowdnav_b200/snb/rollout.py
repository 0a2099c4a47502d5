"""Episode rollouts of B environments to done / time limit: the test loop of the reference's simple_test.py:216-269 (reset a test case,
`while not done: action = robot.act(ob); ob, _, done, info = env.step(action)`, count collisions / wall collisions / frozen / too-close
steps, summarise :306-319) for a whole shard of test cases at once, sharded over ranks by global case id (BASELINE configs[4]).

Everything inside the loop is libsnb launches on one stream: robot policy, (optionally) the JMID prediction + MPC ingest that
SICNavAcados.predict runs before every action (sicnav_acados.py:1640-1667), the fused env step, the metric update.  The only host
synchronisation is the "all done?" poll every `poll_every` steps.
"""
import ctypes as C
import math

import torch

from . import _capi
from .dist import EpisodeMetrics, gather_metrics, global_case_ids, summarize


class LinearRobot:
    """Device stand-in for the robot policy: Linear (crowd_sim_plus/envs/policy/linear.py:16-23) for every environment."""

    kinematics = "holonomic"

    def __init__(self, env):
        self.env = env
        self.action = torch.zeros(env.B, 2, dtype=torch.float64, device=env.device)

    def act(self, stream=None):
        st = self.env.state.cstruct()
        _capi.check(_capi.lib.snb_robot_linear_action(C.byref(st), float(self.env.robot_v_pref), _capi.ptr(self.action),
                                                      _capi.stream_ptr(stream)), "snb_robot_linear_action")
        return self.action


def run_episodes(env, cases, robot=None, forecaster=None, horiz=4, max_steps=None, poll_every=8, phase="test", on_step=None):
    """Resets `env` to the given test cases and steps every environment until it is done (goal reached or time limit,
    crowd_sim_plus.py:1090-1114).  Returns (EpisodeMetrics, steps run, env-steps advanced, last ingest outputs or None).
    forecaster: a ForecasterBatch; when given, the history rings are bootstrapped from the environment's state log
    (sicnav_acados.py:1163-1182) and every step runs update_state_hists + predict_ret_best + the MPC ingest before the robot acts."""
    env.freeze_done = True
    robot = robot or LinearRobot(env)
    env.set_robot_kinematics(getattr(robot, "kinematics", "holonomic"))
    env.reset(phase, test_cases=cases)
    em = EpisodeMetrics(env.B, env.device, env.time_step)
    if max_steps is None:
        max_steps = int(math.ceil(env.time_limit / env.time_step)) + 2
    st = env.state
    ingest = None
    if forecaster is not None:
        forecaster.reset_history()
        if env.n_logged >= env.LOG_DEPTH:
            forecaster.bootstrap_history(env.state_log, (env.n_logged - 1) % env.LOG_DEPTH)
    steps, env_steps = 0, 0
    n_live = env.B
    while steps < max_steps and n_live > 0:
        if forecaster is not None:
            forecaster.push(st.px, st.py, st.rpx, st.rpy)                      # update_state_hists(state, global_time)
            fc, lw = forecaster.predict(env.B)                                 # predict_ret_best
            ingest = forecaster.ingest(fc, lw, horiz=horiz)
        env.step(robot.act())
        em.update(env.flags, env.dmin)
        if on_step is not None:
            on_step(steps, env)
        steps += 1
        env_steps += n_live                      # upper bound between polls; corrected below from the step counters
        if steps % poll_every == 0 or steps >= max_steps:
            n_live = int(env.active.sum().item())
    env.check_status()
    return em, steps, int(em.m[:, 2].sum().item()), ingest


def run_sharded(env_factory, total_envs, rank=0, world=1, test_size=500, **kw):
    """configs[4]: `total_envs` episodes sharded over `world` ranks by global environment id (case = id % test_size, so the set of
    episodes does not depend on the number of ranks); one NCCL all_gather of the [B,9] metric matrix at the end.
    env_factory(n_envs) -> a configured CrowdSimPlusBatch.  Returns (summary dict over ALL environments, gathered metrics, local stats)."""
    cases = global_case_ids(total_envs, rank, world, test_size)
    env = env_factory(len(cases))
    em, steps, env_steps, _ = run_episodes(env, cases, **kw)
    allm = gather_metrics(em.m, total_envs=total_envs)
    return summarize(allm), allm, dict(steps=steps, env_steps=env_steps, envs=len(cases))
