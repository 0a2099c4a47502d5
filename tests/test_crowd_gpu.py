"""GPU parity tests of the crowd step (run with -m gpu on the B200 box).  Everything goes through the C ABI of
libsnb.so (ctypes, snb._capi); the checker is the CPU oracle (oracle/liboracle.so) and the reference-generated
golden rollouts under tests/golden/.

Tolerances (stated per north_star):
  * ORCA / ORCAPlus velocities and RVO2 neighbour lists: BIT-EXACT against the float32 oracle.
  * ORCA rollouts (fp64 positions, bit-exact fp32 velocities): <= 1e-12 m against the oracle, <= 1e-9 against
    the reference-generated golden episodes.
  * SFM (fp64, warp-shuffle summation order + CUDA libm exp differ from numpy by rounding): <= 1e-11 relative per
    step, <= 1e-8 m over the golden episodes.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from golden_util import HCOLS, door_params, load_rollout, rollout_files

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _snb():
    from snb import _capi, state
    from snb.policy import _device_policy
    return _capi, state, _device_policy


def _random_env(rng, B, H, segs=None, spread=4.0, speed=1.0, rradius=0.25, hradius=0.3):
    env = ol.EnvArrays(B, H, segs, rradius=rradius)
    n = B * H
    env.px[:] = rng.uniform(-spread, spread, n); env.py[:] = rng.uniform(-spread, spread, n)
    env.vx[:] = rng.uniform(-speed, speed, n); env.vy[:] = rng.uniform(-speed, speed, n)
    env.gx[:] = rng.uniform(-spread, spread, n); env.gy[:] = rng.uniform(-spread, spread, n)
    env.fgx[:] = env.gx; env.fgy[:] = env.gy
    env.vpref[:] = rng.uniform(0.5, 1.5, n); env.radius[:] = hradius
    env.theta[:] = np.arctan2(env.vy, env.vx)
    env.rpx[:] = rng.uniform(-spread, spread, B); env.rpy[:] = rng.uniform(-spread, spread, B)
    env.rvx[:] = rng.uniform(-1, 1, B); env.rvy[:] = rng.uniform(-1, 1, B)
    env.rgx[:] = 0.0; env.rgy[:] = spread; env.rtheta[:] = np.arctan2(env.rvy, env.rvx)
    return env


def _upload(env, visible=True, kin=0):
    _capi, state, _ = _snb()
    soa = state.CrowdStateSoA(env.B, env.H, 1, "cuda", kin, visible)
    sh = (env.B, env.H)
    soa.load_numpy(**{n: getattr(env, n).reshape(sh) for n in ol.EnvArrays.HUMAN})
    soa.load_numpy(ex_px=env.rpx.reshape(-1, 1), ex_py=env.rpy.reshape(-1, 1), ex_vx=env.rvx.reshape(-1, 1),
                   ex_vy=env.rvy.reshape(-1, 1), ex_radius=np.full((env.B, 1), env.rradius),
                   rtheta=env.rtheta, rgx=env.rgx, rgy=env.rgy, global_time=env.global_time, prev_dist=env.prev_dist)
    return soa


def _pcfg_pair(policy, **kw):
    """(oracle cfg, snb cfg) with identical numbers."""
    _capi, _, _ = _snb()
    o = ol.default_policy_cfg(policy, **kw)
    s = _capi.PolicyCfg(policy=o.policy, max_neighbors=o.max_neighbors, time_step=o.time_step, neighbor_dist=o.neighbor_dist,
                        time_horizon=o.time_horizon, time_horizon_obst=o.time_horizon_obst, policy_radius=o.policy_radius,
                        max_speed=o.max_speed, safety_space=o.safety_space, sfm_radius=o.sfm_radius, A=o.A, B=o.B, KI=o.KI,
                        A_static=o.A_static, B_static=o.B_static, A_bottleneck=o.A_bottleneck, B_bottleneck=o.B_bottleneck,
                        is_bottleneck=o.is_bottleneck)
    return o, s


def _gpu_policy(scfg, soa, segs=None, want_nbr=True):
    _capi, state, dp = _snb()
    obs = state.Obstacles(segs) if segs is not None and len(segs) else None
    if want_nbr:
        v, nbr, cnt, status = dp.step_batch(scfg, soa, obs, True)
        torch.cuda.synchronize()
        return v.cpu().numpy(), nbr.cpu().numpy(), cnt.cpu().numpy(), int(status.item())
    v, status = dp.step_batch(scfg, soa, obs, False)
    torch.cuda.synchronize()
    return v.cpu().numpy(), None, None, int(status.item())


HALLWAY = np.array([[-0.875, -4, -0.875, 4], [0.875, -4, 0.875, 4]], np.float64)
BOTTLENECK = np.array([[-1.0, -4, -1.0, 4], [1.0, -4, 1.0, 4], [-1.0, 0, -0.5, 0], [0.5, 0, 1.0, 0]], np.float64)
SQUEEZE = np.array([[-1.0, -3.75, -0.5, 0], [-0.5, 0, -1.0, 3.75], [1.0, -3.75, 0.5, 0], [0.5, 0, 1.0, 3.75]], np.float64)


@pytest.mark.parametrize("B,H,spread", [(1024, 10, 4.0), (256, 10, 1.2), (64, 5, 2.0), (7, 3, 1.0), (33, 20, 3.0), (1, 1, 1.0)])
def test_orca_policy_bit_exact(B, H, spread):
    """C2 shape (1024 x 10) + dense / ragged / tiny batches: fp32 velocities and neighbour ids identical to the oracle."""
    rng = np.random.default_rng(100 + B + H)
    env = _random_env(rng, B, H, spread=spread)
    ocfg, scfg = _pcfg_pair("orca")
    v_ref, nbr_ref, cnt_ref = ol.policy_batch(ocfg, env, n_threads=8)
    v, nbr, cnt, status = _gpu_policy(scfg, _upload(env))
    assert status == 0
    assert np.array_equal(cnt, cnt_ref)
    assert np.array_equal(nbr, nbr_ref), "RVO2 neighbour order differs"
    assert np.array_equal(v, v_ref), f"max |dv| = {np.max(np.abs(v - v_ref))}"


def test_orca_exact_ties_follow_kdtree_visit_order():
    """11 agents (H=10 + robot) => RVO2's agent kd-tree splits once; humans on a symmetric lattice give exactly equal
    distances, whose order in the neighbour list is the tree's visit order."""
    B, H = 16, 10
    rng = np.random.default_rng(5)
    env = _random_env(rng, B, H, spread=3.0)
    lattice = np.array([(x, y) for y in (-1.0, 0.0, 1.0) for x in (-1.5, -0.5, 0.5, 1.5)])[:H]
    for b in range(B):
        perm = rng.permutation(H)
        env.px[b * H:(b + 1) * H] = lattice[perm, 0] * (1 + b % 3)
        env.py[b * H:(b + 1) * H] = lattice[perm, 1] * (1 + b % 2)
        env.rpx[b] = 0.0; env.rpy[b] = 0.5 * (b % 4)
    ocfg, scfg = _pcfg_pair("orca")
    v_ref, nbr_ref, cnt_ref = ol.policy_batch(ocfg, env, n_threads=4)
    v, nbr, cnt, status = _gpu_policy(scfg, _upload(env))
    assert status == 0 and np.array_equal(cnt, cnt_ref)
    assert np.array_equal(nbr, nbr_ref)
    assert np.array_equal(v, v_ref)


@pytest.mark.parametrize("name,segs", [("hallway", HALLWAY), ("bottleneck", BOTTLENECK), ("squeeze", SQUEEZE)])
def test_orca_plus_obstacles_bit_exact(name, segs):
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    B, H = 512, 6
    env = _random_env(rng, B, H, segs, spread=0.8, hradius=0.2)
    env.py[:] = rng.uniform(-3, 3, B * H)
    ocfg, scfg = _pcfg_pair("orca_plus", safety_space=0.05)
    v_ref, nbr_ref, cnt_ref = ol.policy_batch(ocfg, env, n_threads=8)
    v, nbr, cnt, status = _gpu_policy(scfg, _upload(env), segs)
    assert status == 0
    assert np.array_equal(nbr, nbr_ref) and np.array_equal(cnt, cnt_ref)
    bad = np.argwhere(v != v_ref)
    assert len(bad) == 0, f"{len(bad)} mismatching components, first {bad[:3]}, max |dv| {np.max(np.abs(v - v_ref))}"


def test_obstacle_bsp_split_matches_oracle():
    """processObstacles splits edges that straddle a split line; the vertex list must equal the oracle's."""
    _capi, state, _ = _snb()
    for segs in (HALLWAY, BOTTLENECK, SQUEEZE):
        obs = state.Obstacles(segs)
        L = ol.lib()
        sim = L.rvo_create(0.25, 10, 10, 2.0, 0.5, 0.3, 1.0, 0, 0)
        for s in segs:
            arr = (C.c_float * 4)(*[float(x) for x in s])
            L.rvo_add_obstacle(sim, arr, 2)
        L.rvo_process_obstacles(sim)
        n = L.rvo_get_num_obstacle_vertices(sim)
        ref = np.zeros((n, 7), np.float32)
        buf = (C.c_float * 7)()
        for i in range(n):
            L.rvo_get_obstacle_vertex(sim, i, buf)
            ref[i] = list(buf)
        L.rvo_destroy(sim)
        assert np.array_equal(obs.vertices(), ref)


@pytest.mark.parametrize("B,H,segs,bottleneck", [(4096, 25, HALLWAY, 0), (512, 6, BOTTLENECK, 1), (64, 3, None, 0), (5, 31, SQUEEZE, 0)])
def test_sfm_policy_matches_oracle(B, H, segs, bottleneck):
    """C3 shape (4096 x 25 + walls) and smaller: fp64 SFM within 1e-11 relative of the oracle."""
    rng = np.random.default_rng(B + H)
    env = _random_env(rng, B, H, segs, spread=0.8 if segs is not None else 3.0, hradius=0.2)
    if segs is not None:
        env.py[:] = rng.uniform(-3.5, 3.5, B * H)
    ocfg, scfg = _pcfg_pair("sfm", is_bottleneck=bottleneck)
    v_ref, _, _ = ol.policy_batch(ocfg, env, n_threads=8, want_nbr=False)
    v, _, _, status = _gpu_policy(scfg, _upload(env), segs, want_nbr=False)
    assert status == 0
    err = np.abs(v - v_ref) / np.maximum(1.0, np.abs(v_ref))
    assert err.max() < 1e-11, err.max()


def _gpu_env_from_golden(g):
    _capi, state, _ = _snb()
    H = int(g["H"])
    kin = _capi.KIN_UNICYCLE if bool(g["unicycle"]) else _capi.KIN_HOLONOMIC
    soa = state.CrowdStateSoA(1, H, 1, "cuda", kin, True)
    soa.load_numpy(**{n: g["h0"][:, j].reshape(1, H) for j, n in enumerate(HCOLS)})
    soa.load_numpy(human_time=g["human_times0"].reshape(1, H))
    r0 = g["r0"]
    soa.load_numpy(ex_px=[[r0[0]]], ex_py=[[r0[1]]], ex_vx=[[r0[2]]], ex_vy=[[r0[3]]], ex_radius=[[float(g["robot_radius"])]],
                   rtheta=[r0[4]], rgx=[r0[5]], rgy=[r0[6]], global_time=[float(g["global_time0"])])
    _, scfg = _pcfg_pair(g["human_policy"], time_step=float(g["time_step"]), safety_space=float(g["safety_space"]),
                         sfm_radius=float(g["policy_radius"]), is_bottleneck=int(g["is_bottleneck"]))
    d = door_params(g)
    door = _capi.DoorCfg(enabled=d[0], door_y_mid_min=d[1], door_y_mid_max=d[2], door_x_mid=d[3], door_y_min=d[4],
                         door_y_max=d[5], door_width=d[6])
    o = ol.default_reward_cfg(time_limit=float(g["time_limit"]))
    rcfg = _capi.RewardCfg(success_reward=o.success_reward, timeout=o.timeout, collision_penalty=o.collision_penalty,
                           wall_collision_penalty=o.wall_collision_penalty, freezing_penalty=o.freezing_penalty,
                           discomfort=o.discomfort, has_progress=0, discomfort_dist=o.discomfort_dist,
                           discomfort_penalty_factor=o.discomfort_penalty_factor, progress_factor=0.0, time_limit=o.time_limit)
    obs = state.Obstacles(g["segs"]) if len(g["segs"]) else None
    return soa, scfg, door, rcfg, obs


@pytest.mark.parametrize("path", rollout_files(), ids=lambda p: p.split("rollout_")[-1][:-4])
def test_env_step_replays_reference_episode(path):
    """Step-by-step replay of an episode recorded from the REFERENCE CrowdSimPlus (oracle/gen_golden.py) through
    snb_env_step: states, reward, flags, dmin, human arrival times, clock."""
    _capi, state, _ = _snb()
    g = load_rollout(path)
    soa, scfg, door, rcfg, obs = _gpu_env_from_golden(g)
    tol = 1e-9 if g["human_policy"] != "sfm" else 1e-8
    reward = torch.zeros(1, dtype=torch.float64, device="cuda")
    dmin = torch.zeros(1, dtype=torch.float64, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    for k in range(len(g["reward"])):
        act = torch.tensor(g["actions"][k:k + 1], dtype=torch.float64, device="cuda")
        st = soa.cstruct()
        _capi.check(_capi.lib.snb_env_step(C.byref(scfg), C.byref(door), C.byref(rcfg), C.byref(st),
                                           obs.handle if obs is not None else None, _capi.ptr(act), None, _capi.ptr(reward),
                                           _capi.ptr(dmin), _capi.ptr(flags), None, None, _capi.ptr(status),
                                           _capi.stream_ptr()), "snb_env_step")
        torch.cuda.synchronize()
        hs = np.stack([getattr(soa, n).cpu().numpy()[0] for n in HCOLS], 1)
        rs = np.array([soa.ex_px[0, 0].item(), soa.ex_py[0, 0].item(), soa.ex_vx[0, 0].item(), soa.ex_vy[0, 0].item(),
                       soa.rtheta[0].item(), soa.rgx[0].item(), soa.rgy[0].item()])
        assert np.max(np.abs(hs - g["H_states"][k])) < tol, (k, np.max(np.abs(hs - g["H_states"][k])))
        assert np.max(np.abs(rs - g["R_states"][k])) < tol, k
        assert abs(reward.item() - g["reward"][k]) < 1e-8, (k, reward.item(), g["reward"][k])
        assert int(flags.item()) == int(g["flags"][k]), (k, int(flags.item()), int(g["flags"][k]))
        if not np.isnan(g["dmin"][k]):
            assert abs(dmin.item() - g["dmin"][k]) < 1e-8
        assert np.max(np.abs(soa.human_time.cpu().numpy()[0] - g["human_times"][k])) < 1e-9, k
        assert abs(soa.global_time[0].item() - g["global_time"][k]) < 1e-12
    assert int(status.item()) == 0


@pytest.mark.parametrize("policy,B,H,segs,steps", [("orca", 1024, 10, None, 40), ("orca_plus", 256, 6, BOTTLENECK, 25),
                                                     ("sfm", 1024, 25, HALLWAY, 15)])
def test_env_step_batch_rollout_vs_oracle(policy, B, H, segs, steps):
    """Multi-step batched rollout (in-place state in HBM) vs the oracle stepping the same scenes on the CPU.
    ORCA: bit-identical trajectories.  SFM: <= 1e-8 m drift."""
    _capi, state, _ = _snb()
    rng = np.random.default_rng(42)
    env = _random_env(rng, B, H, segs, spread=3.0 if segs is None else 0.8, speed=0.5, hradius=0.3 if segs is None else 0.2)
    if segs is not None:
        env.py[:] = rng.uniform(-3.5, 3.5, B * H)
        env.rpx[:] = rng.uniform(-0.5, 0.5, B)
    ocfg, scfg = _pcfg_pair(policy, safety_space=0.05 if policy == "orca_plus" else 0.0)
    orc = ol.default_reward_cfg(time_limit=2.0)
    rcfg = _capi.RewardCfg(success_reward=orc.success_reward, timeout=orc.timeout, collision_penalty=orc.collision_penalty,
                           wall_collision_penalty=orc.wall_collision_penalty, freezing_penalty=orc.freezing_penalty,
                           discomfort=orc.discomfort, has_progress=0, discomfort_dist=orc.discomfort_dist,
                           discomfort_penalty_factor=orc.discomfort_penalty_factor, progress_factor=0.0, time_limit=orc.time_limit)
    door = _capi.DoorCfg(enabled=0)
    odoor = ol.DoorCfg(enabled=0)
    soa = _upload(env)
    obs = state.Obstacles(segs) if segs is not None else None
    reward = torch.zeros(B, dtype=torch.float64, device="cuda")
    dmin = torch.zeros(B, dtype=torch.float64, device="cuda")
    flags = torch.zeros(B, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    for k in range(steps):
        act = rng.uniform(-1, 1, (B, 2))
        r_ref, d_ref, f_ref = ol.env_step(ocfg, odoor, orc, env, act, n_threads=8)
        a = torch.tensor(act, dtype=torch.float64, device="cuda")
        st = soa.cstruct()
        _capi.check(_capi.lib.snb_env_step(C.byref(scfg), C.byref(door), C.byref(rcfg), C.byref(st),
                                           obs.handle if obs is not None else None, _capi.ptr(a), None, _capi.ptr(reward),
                                           _capi.ptr(dmin), _capi.ptr(flags), None, None, _capi.ptr(status), _capi.stream_ptr()),
                    "snb_env_step")
        torch.cuda.synchronize()
        got = soa.to_numpy("px", "py", "vx", "vy", "theta", "gx", "gy", "human_time", "ex_px", "ex_py", "ex_vx", "ex_vy",
                           "rtheta", "global_time")
        if policy != "sfm":
            for n in ("px", "py", "vx", "vy"):
                assert np.array_equal(got[n].ravel(), getattr(env, n)), (k, n, np.max(np.abs(got[n].ravel() - getattr(env, n))))
            assert np.array_equal(flags.cpu().numpy(), f_ref), k
        else:
            for n in ("px", "py", "vx", "vy"):
                assert np.max(np.abs(got[n].ravel() - getattr(env, n))) < 1e-8, (k, n)
            assert np.mean(flags.cpu().numpy() == f_ref) > 0.999
        assert np.max(np.abs(got["ex_px"].ravel() - env.rpx)) < 1e-12 and np.max(np.abs(got["global_time"] - env.global_time)) == 0
        assert np.max(np.abs(reward.cpu().numpy() - r_ref)) < (1e-12 if policy != "sfm" else 1e-7)
    assert int(status.item()) == 0


@pytest.mark.parametrize("policy,B,H,segs,kin", [("orca", 512, 10, None, 0), ("orca_plus", 128, 6, BOTTLENECK, 1), ("sfm", 256, 25, HALLWAY, 0)])
def test_env_whatif_equals_step_on_a_copy(policy, B, H, segs, kin):
    """snb_env_whatif = CrowdSimPlus.step(action, update=False) for A candidate robot actions per environment
    (crowd_sim_plus.py:1239-1255; the RL look-ahead of :797-866).  Oracle: the C restatement of step() applied to a COPY of the
    scene, once per candidate action.  ORCA: reward / flags / next states bit-identical; SFM: 1e-8.  The state in HBM and
    prev_dist must not change."""
    _capi, state, _ = _snb()
    rng = np.random.default_rng(7)
    A = 31                                                         # 1 + 30 discrete actions (build_action_space, :275-297)
    env = _random_env(rng, B, H, segs, spread=3.0 if segs is None else 0.8, speed=0.5, hradius=0.3 if segs is None else 0.2)
    env.robot_kinematics = kin
    if segs is not None:
        env.py[:] = rng.uniform(-3.5, 3.5, B * H)
        env.rpx[:] = rng.uniform(-0.5, 0.5, B)
    env.prev_dist[:] = np.hypot(env.rpx - env.rgx, env.rpy - env.rgy)
    ocfg, scfg = _pcfg_pair(policy, safety_space=0.05 if policy == "orca_plus" else 0.0)
    orc = ol.default_reward_cfg(time_limit=2.0)
    orc.has_progress = 1; orc.progress_factor = 0.3
    rcfg = _capi.RewardCfg(success_reward=orc.success_reward, timeout=orc.timeout, collision_penalty=orc.collision_penalty,
                           wall_collision_penalty=orc.wall_collision_penalty, freezing_penalty=orc.freezing_penalty,
                           discomfort=orc.discomfort, has_progress=1, discomfort_dist=orc.discomfort_dist,
                           discomfort_penalty_factor=orc.discomfort_penalty_factor, progress_factor=0.3, time_limit=orc.time_limit)
    door, odoor = _capi.DoorCfg(enabled=0), ol.DoorCfg(enabled=0)
    soa = _upload(env, kin=kin)
    obs = state.Obstacles(segs) if segs is not None else None
    acts = rng.uniform(-1, 1, (B, A, 2))
    acts[:, 0] = 0.0                                                # ActionRot(0, 0): frozen
    before = soa.to_numpy("px", "py", "vx", "vy", "gx", "gy", "ex_px", "ex_py", "rtheta", "global_time", "prev_dist")
    a = torch.tensor(acts, dtype=torch.float64, device="cuda")
    reward = torch.zeros(B, A, dtype=torch.float64, device="cuda"); dmin = torch.zeros_like(reward)
    flags = torch.zeros(B, A, dtype=torch.int32, device="cuda")
    nh = torch.zeros(B, H, 4, dtype=torch.float64, device="cuda"); nr = torch.zeros(B, A, 2, dtype=torch.float64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = soa.cstruct()
    _capi.check(_capi.lib.snb_env_whatif(C.byref(scfg), C.byref(door), C.byref(rcfg), C.byref(st), obs.handle if obs is not None else None,
                                         _capi.ptr(a), A, None, _capi.ptr(reward), _capi.ptr(dmin), _capi.ptr(flags), _capi.ptr(nh),
                                         _capi.ptr(nr), _capi.ptr(status), _capi.stream_ptr()), "snb_env_whatif")
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    after = soa.to_numpy("px", "py", "vx", "vy", "gx", "gy", "ex_px", "ex_py", "rtheta", "global_time", "prev_dist")
    for n in before:
        assert np.array_equal(before[n], after[n]), f"what-if modified state.{n}"
    tol = 0.0 if policy != "sfm" else 1e-8
    R, F, NH, NR = reward.cpu().numpy(), flags.cpu().numpy(), nh.cpu().numpy(), nr.cpu().numpy()
    for k in range(A):
        e2 = env.copy()
        r_ref, d_ref, f_ref = ol.env_step(ocfg, odoor, orc, e2, np.ascontiguousarray(acts[:, k]), n_threads=8)
        assert np.max(np.abs(NR[:, k, 0] - e2.rpx)) <= 1e-12 and np.max(np.abs(NR[:, k, 1] - e2.rpy)) <= 1e-12, k
        assert np.max(np.abs(R[:, k] - r_ref)) <= (1e-12 if policy != "sfm" else 1e-7), k
        if policy != "sfm":
            assert np.array_equal(F[:, k], f_ref), k
        else:
            assert np.mean(F[:, k] == f_ref) > 0.995, k
        if k == 0:
            ref_nh = np.stack([e2.px, e2.py, e2.vx, e2.vy], -1).reshape(B, H, 4)
            assert np.max(np.abs(NH - ref_nh)) <= tol
    assert (F[:, 0] & _capi.F_FROZEN).all()


def test_policy_plugin_predict_matches_oracle():
    """The B=1 plugin call: snb.policy.ORCA().predict(JointState) == what the reference's orca.py computes on the
    oracle RVO2, including the neighbour order; SFM().predict within 1e-12."""
    from snb.policy import ORCA, SFM, ORCAPlus, policy_factory
    from snb.utils.state_plus import FullState, JointState, ObservableState
    rng = np.random.default_rng(3)
    L = ol.lib()
    for trial in range(40):
        n = int(rng.integers(0, 12))
        me = FullState(*rng.uniform(-2, 2, 2), *rng.uniform(-1, 1, 2), 0.3, *rng.uniform(-4, 4, 2), float(rng.uniform(0.5, 1.5)), 0.0)
        others = [ObservableState(*rng.uniform(-3, 3, 2), *rng.uniform(-1, 1, 2), 0.3) for _ in range(n)]
        segs = [[(-0.875, -4.0), (-0.875, 4.0)], [(0.875, -4.0), (0.875, 4.0)]] if trial % 2 else []
        self8 = np.array([me.px, me.py, me.vx, me.vy, me.radius, me.gx, me.gy, me.v_pref])
        oth = np.array([o.as_tuple() for o in others], np.float64).ravel()
        sg = np.array(segs, np.float64).ravel()
        for cls, kind in ((ORCA, "orca"), (ORCAPlus, "orca_plus"), (SFM, "sfm")):
            pol = policy_factory[kind]()
            assert isinstance(pol, cls)
            pol.time_step = 0.25
            if kind == "sfm":
                for k_, v_ in dict(radius=0.2, A=3.0, B=0.18, KI=1.0, A_static=2.0, B_static=0.025, A_bottleneck=6.0, B_bottleneck=0.12).items():
                    setattr(pol, k_, v_)
            if kind == "orca_plus":
                pol.safety_space = 0.05
            act = pol.predict(JointState(me, others, segs))
            ocfg = ol.default_policy_cfg(kind, safety_space=0.05 if kind == "orca_plus" else 0.0)
            out = np.zeros(2)
            if kind == "sfm":
                L.orc_sfm_predict(C.byref(ocfg), ol.dptr(self8), n, ol.dptr(oth) if n else None, len(sg) // 4,
                                  ol.dptr(sg) if len(sg) else None, ol.dptr(out))
                assert np.allclose([act.vx, act.vy], out, rtol=1e-12, atol=1e-12)
            else:
                ids = np.zeros(32, np.int32); cnt = C.c_int(0)
                L.orc_orca_predict(C.byref(ocfg), ol.dptr(self8), n, ol.dptr(oth) if n else None, len(sg) // 4,
                                   ol.dptr(sg) if len(sg) else None, ol.dptr(out), ol.iptr(ids), C.byref(cnt), None, None)
                assert (act.vx, act.vy) == (out[0], out[1]), (kind, trial)
                assert pol.last_neighbors == list(ids[:cnt.value])


def test_crowdsimplus_batch_matches_reference_reset_and_episode():
    """CrowdSimPlusBatch.configure/reset/step against a reference episode: same seeded scene, same trajectory."""
    import configparser
    from snb.env import CrowdSimPlusBatch
    g = load_rollout([p for p in rollout_files() if p.endswith("orca_circle10_warm.npz")][0])
    cfg = configparser.RawConfigParser()
    cfg.read_string("""
[env]
time_limit = 30
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
human_num = 10
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
""")
    env = CrowdSimPlusBatch(3, "cuda")
    env.configure(cfg)
    env.reset('test', test_cases=[1, 1, 0])
    h0 = np.stack([getattr(env.state, n).cpu().numpy()[0] for n in HCOLS], 1)
    assert np.max(np.abs(h0 - g["h0"])) < 1e-9          # after the 10 warm-up steps
    for k in range(len(g["reward"])):
        a = np.tile(g["actions"][k], (3, 1))
        reward, done, flags = env.step(a)
        hs = np.stack([getattr(env.state, n).cpu().numpy()[0] for n in HCOLS], 1)
        assert np.max(np.abs(hs - g["H_states"][k])) < 1e-9, k
        assert abs(reward[0].item() - g["reward"][k]) < 1e-9
        assert int(flags[0].item()) == int(g["flags"][k])
    env.check_status()


class _CountingRng:
    """numpy Generator wrapper that counts the 64-bit draws the host generator consumes."""
    def __init__(self, seed):
        self.g, self.n = np.random.default_rng(seed), 0

    def random(self):
        self.n += 1
        return self.g.random()

    def uniform(self, a, b):
        self.n += 1
        return self.g.uniform(a, b)


@pytest.mark.parametrize("rule,H,B", [("circle_crossing", 10, 1024), ("circle_crossing", 5, 64), ("hallway", 6, 256),
                                      ("hallway_static", 5, 256), ("hallway_bottleneck", 5, 256), ("hallway_squeeze", 4, 128),
                                      ("rectangle", 6, 128)])
def test_scene_reset_on_device_matches_host_generator(rule, H, B):
    """snb_scene_reset (SeedSequence + PCG64 + the reference's rejection sampling, one thread per environment) against the host
    restatement of CrowdSimPlus.reset (oracle/scenario_oracle.py, pinned to reference episodes): same number of draws per environment (=
    identical accept / reject sequence), v_pref bit-equal (pure PCG64 arithmetic), positions / goals within 1e-12 (libm ulps)."""
    import configparser
    import scenario_oracle as scenario
    from snb.env import CrowdSimPlusBatch
    cfg = configparser.RawConfigParser()
    cfg.read_string(f"""
[env]
time_limit = 30
time_step = 0.25
val_size = 100
test_size = 500
randomize_attributes = true
[sim]
train_val_sim = {rule}
test_sim = {rule}
starts_moving = 0
square_width = 5
circle_radius = 4.0
rect_width = 2.5
rect_height = 4
human_num = {H}
[humans]
visible = true
policy = {'orca' if rule == 'circle_crossing' else 'orca_plus'}
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
""")
    env = CrowdSimPlusBatch(B, "cuda")
    env.configure(cfg)
    cases = (np.arange(B) * 7 + 3) % 500
    env.reset('test', test_cases=cases)
    torch.cuda.synchronize()
    dev = env.state.to_numpy("px", "py", "gx", "gy", "fgx", "fgy", "vpref", "theta", "radius", "ex_px", "ex_py", "rgx", "rgy", "rtheta", "prev_dist")
    draws = env.reset_draws.cpu().numpy()
    p = scenario.SceneParams(env.circle_radius, env.rect_width, env.rect_height, env.human_radius, env.human_v_pref,
                             env.robot_radius, env.rewards["discomfort_dist"], env.randomize_attributes)
    segs, door = scenario.static_obstacles(rule, p)
    for b in range(B):
        rng = _CountingRng(1000 + int(cases[b]))          # test phase: offset = case_capacity['val'] = 1000
        raw = scenario._generate_with_goals(rule, H, rng, p, segs, door)
        assert rng.n == draws[b], (b, rng.n, draws[b])
        for i, (px, py, fgx, fgy, v_pref, theta) in enumerate(raw):
            gx, gy = scenario.door_goal(rule, door, len(segs), px, py, fgx, fgy)
            got = [dev[n][b, i] for n in ("px", "py", "gx", "gy", "fgx", "fgy", "theta")]
            assert np.max(np.abs(np.array(got) - np.array([px, py, gx, gy, fgx, fgy, theta]))) < 1e-12, (b, i)
            assert dev["vpref"][b, i] == v_pref
    assert np.all(dev["radius"] == 0.3) and np.all(dev["ex_py"] == -4.0) and np.all(dev["rgy"] == 4.0) and np.all(dev["prev_dist"] == 8.0)
    # and the simulator runs from the device-built scenes exactly as from the uploaded ones
    env2 = CrowdSimPlusBatch(B, "cuda")
    env2.configure(cfg)
    env2.reset('test', test_cases=cases)
    hum = np.stack([scenario.generate_scene(rule, H, int(c), 'test', p)["humans"] for c in cases])     # host-generated scenes, uploaded
    env2.state.load_numpy(px=hum[..., 0], py=hum[..., 1], gx=hum[..., 2], gy=hum[..., 3], fgx=hum[..., 4], fgy=hum[..., 5],
                          vpref=hum[..., 6], theta=hum[..., 7])
    a = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); a[:, 1] = 1.0
    for _ in range(5):
        env.step(a); env2.step(a)
    assert float((env.state.px - env2.state.px).abs().max()) < 1e-6


def test_scene_reset_gives_up_loudly_on_an_overcrowded_scene():
    """25 humans cannot be placed on a 4 m circle with 0.8 m clearance from every position AND goal: the reference's rejection
    sampling spins forever; snb_scene_reset stops after 200 000 tries per human, reports -1 draws and reset() raises."""
    import configparser
    from snb import _capi
    from snb.env import CrowdSimPlusBatch
    cfg = configparser.RawConfigParser()
    cfg.read_string("""
[env]
time_limit = 30
time_step = 0.25
val_size = 100
test_size = 500
randomize_attributes = true
[sim]
train_val_sim = circle_crossing
test_sim = circle_crossing
starts_moving = 0
square_width = 5
circle_radius = 4.0
rect_width = 2.5
rect_height = 4
human_num = 25
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
""")
    env = CrowdSimPlusBatch(4, "cuda")
    env.configure(cfg)
    with pytest.raises(_capi.SnbError, match="gave up"):
        env.reset('test', test_cases=[0, 1, 2, 3])
    assert int(env.reset_draws.min().item()) == -1


def test_whatif_through_the_batch_env_matches_a_committed_step():
    """CrowdSimPlusBatch.what_if(actions)[:, k] == what step(actions[:, k]) then returns, for every candidate k (ORCA, bit-exact),
    and the environment can still be stepped afterwards."""
    import configparser
    from snb.env import CrowdSimPlusBatch
    cfg = configparser.RawConfigParser()
    cfg.read_string("""
[env]
time_limit = 30
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
rect_width = 2.5
rect_height = 4
human_num = 8
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
""")
    B, A = 64, 5
    env = CrowdSimPlusBatch(B, "cuda")
    env.configure(cfg)
    env.reset('test')
    rng = np.random.default_rng(3)
    acts = torch.tensor(rng.uniform(-1, 1, (B, A, 2)), dtype=torch.float64, device="cuda")
    reward, done, flags, next_h, next_r = env.what_if(acts)
    torch.cuda.synchronize()
    for k in range(A):
        e2 = CrowdSimPlusBatch(B, "cuda")
        e2.configure(cfg)
        e2.reset('test')
        r2, d2, f2 = e2.step(acts[:, k].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(r2, reward[:, k]) and torch.equal(f2, flags[:, k]) and torch.equal(d2, done[:, k])
        assert torch.equal(e2.state.px, next_h[:, :, 0]) and torch.equal(e2.state.vy, next_h[:, :, 3])
        assert torch.equal(e2.state.rpx, next_r[:, k, 0]) and torch.equal(e2.state.rpy, next_r[:, k, 1])
    env.step(acts[:, 0].contiguous())
    env.check_status()


@pytest.mark.parametrize("case", ["policy_1024x10", "policy_dense", "policy_ragged", "policy_single", "ties", "hallway", "bottleneck", "squeeze",
                                  "rollout_orca", "rollout_orca_plus", "whatif_orca", "whatif_orca_plus", "sfm_4096x25", "sfm_bottleneck",
                                  "rollout_sfm", "whatif_sfm", "sfm_reference_episodes"])
def test_thread_per_human_orca_is_bit_identical(monkeypatch, case):
    """The one-thread-per-human path (large batches) against the same oracle / reference comparisons as the warp-cooperative path:
    ORCA neighbour lists, velocities, multi-step rollouts and the what-if step stay bit-exact with SNB_CROWD_MODE=thread; SFM keeps
    its tolerances (it now sums in the reference's sequential order)."""
    monkeypatch.setenv("SNB_CROWD_MODE", "thread")
    if case == "policy_1024x10":
        test_orca_policy_bit_exact(1024, 10, 4.0)
    elif case == "policy_dense":
        test_orca_policy_bit_exact(256, 10, 1.2)
    elif case == "policy_ragged":
        test_orca_policy_bit_exact(33, 20, 3.0)
    elif case == "policy_single":
        test_orca_policy_bit_exact(1, 1, 1.0)
    elif case == "ties":
        test_orca_exact_ties_follow_kdtree_visit_order()
    elif case in ("hallway", "bottleneck", "squeeze"):
        test_orca_plus_obstacles_bit_exact(case, {"hallway": HALLWAY, "bottleneck": BOTTLENECK, "squeeze": SQUEEZE}[case])
    elif case == "rollout_orca":
        test_env_step_batch_rollout_vs_oracle("orca", 1024, 10, None, 40)
    elif case == "rollout_orca_plus":
        test_env_step_batch_rollout_vs_oracle("orca_plus", 256, 6, BOTTLENECK, 25)
    elif case == "whatif_orca":
        test_env_whatif_equals_step_on_a_copy("orca", 512, 10, None, 0)
    elif case == "whatif_orca_plus":
        test_env_whatif_equals_step_on_a_copy("orca_plus", 128, 6, BOTTLENECK, 1)
    elif case == "sfm_4096x25":
        test_sfm_policy_matches_oracle(4096, 25, HALLWAY, 0)
    elif case == "sfm_bottleneck":
        test_sfm_policy_matches_oracle(512, 6, BOTTLENECK, 1)
    elif case == "rollout_sfm":
        test_env_step_batch_rollout_vs_oracle("sfm", 1024, 25, HALLWAY, 15)
    elif case == "whatif_sfm":
        test_env_whatif_equals_step_on_a_copy("sfm", 256, 25, HALLWAY, 0)
    else:
        for path in rollout_files():
            if "sfm" in os.path.basename(path):
                test_env_step_replays_reference_episode(path)
