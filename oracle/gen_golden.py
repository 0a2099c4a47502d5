"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):  python oracle/gen_golden.py
  * SFM / geometry / clamp / CrowdSimPlus.step: the reference's own Python
    (crowd_sim_plus/envs/...) imported through oracle/ref_shims.py.
  * ORCA humans inside those rollouts: the reference's own orca.py / orca_plus.py running on
    oracle/rvo2_shim (upstream Python-RVO2 is not installable; parity for the RVO2 arithmetic
    itself stays UNPINNED).
  * JMID / iMID denoiser: the reference's models/diffusion.py with (a) the shipped checkpoints,
    (b) seeded synthetic weights (oracle/jmid_oracle.make_random_weights) loaded into it.
"""
import configparser
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

import ref_shims  # noqa: E402

assert ref_shims.have_reference(), "needs /root/reference"
ref_shims.install_crowd_sim_shims()
import torch  # noqa: E402

from crowd_sim_plus.envs.policy.policy import Policy  # noqa: E402
from crowd_sim_plus.envs.policy.social_force import SFM  # noqa: E402
from crowd_sim_plus.envs.utils.action import ActionRot, ActionXY  # noqa: E402
from crowd_sim_plus.envs.utils.state_plus import FullState, JointState, ObservableState  # noqa: E402
from crowd_sim_plus.envs.utils import utils_plus as U  # noqa: E402
import crowd_sim_plus.envs.crowd_sim_plus as csp  # noqa: E402
from crowd_sim_plus.envs.utils.robot_plus import Robot  # noqa: E402

REF_CFG = os.path.join(ref_shims.REF, "sicnav_diffusion/configs/env.config")


def env_config(**over):
    cfg = configparser.RawConfigParser()
    cfg.read(REF_CFG)
    cfg.set("robot", "policy", "linear")
    for k, v in over.items():
        sec, key = k.split("__")
        cfg.set(sec, key, str(v))
    return cfg


# ------------------------------------------------------------------ SFM
def gen_sfm():
    rng = np.random.default_rng(7)
    cfg = env_config()
    cases = []
    for c in range(96):
        s = SFM()
        s.configure(cfg, "humans")
        s.is_bottleneck = bool(c % 3 == 0)
        n = int(rng.integers(0, 26))
        m = int(rng.integers(0, 13))
        me = FullState(*rng.uniform(-2, 2, 2), *rng.uniform(-1, 1, 2), rng.uniform(0.15, 0.35), *rng.uniform(-4, 4, 2),
                       rng.uniform(0.5, 1.5), 0.0)
        if c == 5:  # goal == position branch (dist_to_goal < 1e-6)
            me = FullState(me.px, me.py, me.vx, me.vy, me.radius, me.px, me.py, me.v_pref, 0.0)
        others = [ObservableState(*rng.uniform(-3, 3, 2), *rng.uniform(-1, 1, 2), rng.uniform(0.15, 0.35)) for _ in range(n)]
        segs = []
        for _ in range(m):
            p = rng.uniform(-3, 3, 2)
            q = p + rng.uniform(-3, 3, 2)
            segs.append([(float(p[0]), float(p[1])), (float(q[0]), float(q[1]))])
        a = s.predict(JointState(me, others, segs))
        cases.append(dict(self8=[me.px, me.py, me.vx, me.vy, me.radius, me.gx, me.gy, me.v_pref],
                          others=np.array([[o.px, o.py, o.vx, o.vy, o.radius] for o in others]).reshape(-1, 5),
                          segs=np.array(segs, np.float64).reshape(-1, 4), bottleneck=s.is_bottleneck,
                          out=[float(a.vx), float(a.vy)]))
    # SURVEY Appendix C.3 known answers
    np.savez(os.path.join(OUT, "sfm_cases.npz"),
             n=len(cases),
             **{f"self8_{i}": np.array(c["self8"]) for i, c in enumerate(cases)},
             **{f"others_{i}": c["others"] for i, c in enumerate(cases)},
             **{f"segs_{i}": c["segs"] for i, c in enumerate(cases)},
             bottleneck=np.array([c["bottleneck"] for c in cases]),
             out=np.array([c["out"] for c in cases]),
             sfm_params=np.array([0.2, 3.0, 0.18, 1.0, 2.0, 0.025, 6.0, 0.12, 0.25]))
    print("sfm:", len(cases))


# ------------------------------------------------------------------ geometry + clamp
class _A:  # minimal agent with the attributes constrain_agent_action_exact touches
    def __init__(self, px, py, theta, radius, kin, dt):
        self.px, self.py, self.theta, self.radius, self.kinematics, self.time_step = px, py, theta, radius, kin, dt

    def compute_position(self, action, delta_t):
        if self.kinematics == "holonomic":
            return self.px + action.vx * delta_t, self.py + action.vy * delta_t
        theta = self.theta + action.r
        return self.px + np.cos(theta) * action.v * delta_t, self.py + np.sin(theta) * action.v * delta_t


def gen_geometry():
    rng = np.random.default_rng(11)
    segseg_in, segseg_out = [], []
    for c in range(400):
        a0 = rng.uniform(-2, 2, 2)
        a1 = a0 + rng.uniform(-3, 3, 2)
        b0 = rng.uniform(-2, 2, 2)
        b1 = b0 + rng.uniform(-1, 1, 2)
        k = c % 8
        if k == 1:      # parallel, same direction
            b1 = b0 + (a1 - a0) * rng.uniform(0.1, 0.6)
        elif k == 2:    # parallel, opposite direction, colinear overlap
            b0 = a0 + (a1 - a0) * rng.uniform(0.2, 1.3)
            b1 = b0 - (a1 - a0) * rng.uniform(0.1, 0.9)
        elif k == 3:    # zero-length travel
            b1 = b0.copy()
        elif k == 4:    # axis-aligned wall, parallel travel
            a0 = np.array([-0.875, -4.0]); a1 = np.array([-0.875, 4.0])
            b0 = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-5, 5)]); b1 = b0 + np.array([0.0, rng.uniform(-1, 1)])
        elif k == 5:    # parallel colinear beyond the ends
            b0 = a1 + (a1 - a0) * rng.uniform(0.1, 0.5)
            b1 = b0 + (a1 - a0) * rng.uniform(0.1, 0.5) * (1 if c % 16 < 8 else -0.05)
        try:
            pA, pB, d = U.closest_distance_between_line_segments(np.array([*a0, 0.0]), np.array([*a1, 0.0]),
                                                                 np.array([*b0, 0.0]), np.array([*b1, 0.0]))
        except AssertionError:
            continue
        segseg_in.append([*a0, *a1, *b0, *b1])
        segseg_out.append([pA[0], pA[1], pB[0], pB[1], d])
    # SURVEY Appendix C.4
    pA, pB, d = U.closest_distance_between_line_segments(np.array([-.875, -4, 0.]), np.array([-.875, 4, 0.]),
                                                         np.array([-.7, 0, 0.]), np.array([-.9, .3, 0.]))
    segseg_in.append([-.875, -4, -.875, 4, -.7, 0, -.9, .3]); segseg_out.append([pA[0], pA[1], pB[0], pB[1], d])

    env = csp.CrowdSimPlus()
    layouts = [
        [[(-0.875, -4.0), (-0.875, 4.0)], [(0.875, -4.0), (0.875, 4.0)]],
        [[(-1.0, -4.0), (-1.0, 4.0)], [(1.0, -4.0), (1.0, 4.0)], [(-1.0, 0.0), (-0.5, 0.0)], [(0.5, 0.0), (1.0, 0.0)]],
        [[(-1.0, -2.0), (-0.25, 0.0)], [(-0.25, 0.0), (-1.0, 2.0)], [(1.0, -2.0), (0.25, 0.0)], [(0.25, 0.0), (1.0, 2.0)]],
    ]
    cl_in, cl_out, cl_layout = [], [], []
    for c in range(900):
        li = c % len(layouts)
        env.static_obstacles = layouts[li]
        kin = "holonomic" if c % 3 else "unicycle"
        r = float(rng.uniform(0.2, 0.3))
        # sample near the walls so that the clamp triggers often
        seg = layouts[li][int(rng.integers(len(layouts[li])))]
        t = rng.uniform(-0.1, 1.1)
        base = np.array(seg[0]) + t * (np.array(seg[1]) - np.array(seg[0]))
        nrm = np.array([-(seg[1][1] - seg[0][1]), seg[1][0] - seg[0][0]], float)
        nrm /= np.linalg.norm(nrm)
        off = rng.uniform(r - 1e-5 if c % 5 == 0 else r, r + 0.35) * (1 if rng.random() < 0.5 else -1)
        p = base + nrm * off
        theta = float(rng.uniform(-np.pi, np.pi))
        ag = _A(float(p[0]), float(p[1]), theta, r, kin, 0.25)
        if kin == "holonomic":
            act = ActionXY(*[float(x) for x in rng.uniform(-1.5, 1.5, 2)])
            if c % 7 == 0:
                act = ActionXY(0.0, 0.0)
        else:
            act = ActionRot(float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-0.6, 0.6)))
            if c % 11 == 0:
                act = ActionRot(0.0, act.r)
        try:
            with np.errstate(all="ignore"):
                out = env.constrain_agent_action_exact(ag, act)
        except AssertionError:
            continue
        if not np.all(np.isfinite(np.array(out, float))):
            continue
        cl_in.append([ag.px, ag.py, theta, r, 0.25, 0.0 if kin == "holonomic" else 1.0, act[0], act[1]])
        cl_out.append([float(out[0]), float(out[1])])
        cl_layout.append(li)
    np.savez(os.path.join(OUT, "geometry_cases.npz"), segseg_in=np.array(segseg_in), segseg_out=np.array(segseg_out),
             clamp_in=np.array(cl_in), clamp_out=np.array(cl_out), clamp_layout=np.array(cl_layout),
             **{f"layout_{i}": np.array(l, np.float64).reshape(-1, 4) for i, l in enumerate(layouts)})
    changed = int(np.sum(np.any(np.abs(np.array(cl_in)[:, 6:8] - np.array(cl_out)) > 0, axis=1)))
    print("geometry: segseg", len(segseg_in), "clamp", len(cl_in), "of which clamped", changed)


# ------------------------------------------------------------------ full CrowdSimPlus rollouts
class ScriptedUnicycle(Policy):
    """Robot policy with kinematics='unicycle' emitting ActionRot toward the goal (test driver only)."""

    def __init__(self):
        super().__init__()
        self.name = "scripted"
        self.kinematics = "unicycle"
        self.multiagent_training = True

    def predict(self, state):
        s = state.self_state
        want = np.arctan2(s.gy - s.py, s.gx - s.px)
        d = (want - s.theta + np.pi) % (2 * np.pi) - np.pi
        return ActionRot(0.8 * s.v_pref, float(np.clip(d, -0.5, 0.5)))


class ScriptedHolonomic(Policy):
    """Robot policy emitting a fixed ActionXY pattern (drives into walls / stands still; test driver only)."""

    def __init__(self, pattern):
        super().__init__()
        self.name = "scripted_xy"
        self.kinematics = "holonomic"
        self.multiagent_training = True
        self.pattern = pattern
        self.k = 0

    def predict(self, state):
        a = self.pattern[self.k % len(self.pattern)]
        self.k += 1
        return ActionXY(float(a[0]), float(a[1]))


def rollout(name, human_policy, sim, H, case, steps, starts_moving=0, unicycle=False, pattern=None, **over):
    cfg = env_config(humans__policy=human_policy, sim__test_sim=sim, sim__train_val_sim=sim, sim__human_num=H,
                     sim__starts_moving=starts_moving, **over)
    env = csp.CrowdSimPlus()
    env.configure(cfg)
    robot = Robot(cfg, "robot")
    if unicycle:
        robot.set_policy(ScriptedUnicycle())
    elif pattern is not None:
        robot.set_policy(ScriptedHolonomic(pattern))
    env.set_robot(robot)
    robot.policy.set_env(env)
    ob, static = env.reset("test", case, return_stat=True)

    def snap():
        hs = env.humans
        return (np.array([[h.px, h.py, h.vx, h.vy, h.theta, h.gx, h.gy, h.final_gx, h.final_gy, h.v_pref, h.radius] for h in hs]),
                np.array([robot.px, robot.py, robot.vx, robot.vy, robot.theta, robot.gx, robot.gy]))

    h0, r0 = snap()
    ht0 = np.array(env.human_times, float)
    gt0 = float(env.global_time)
    H_states, R_states, acts, rew, flags, dmins, htimes, gtimes = [], [], [], [], [], [], [], []
    for k in range(steps):
        a = robot.act(ob, static)
        ob, reward, done, info = env.step(a)
        hs, rs = snap()
        H_states.append(hs); R_states.append(rs); acts.append([float(a[0]), float(a[1])]); rew.append(float(reward))
        f = (1 * (info["ReachGoal"].val != 0) | 2 * (info["Timeout"].val != 0) | 4 * (info["Collision"].val != 0)
             | 8 * (info["WallCollision"].val != 0) | 16 * (info["Frozen"].val != 0) | 32 * (info["Danger"].val != 0)
             | 64 * bool(done))
        flags.append(int(f))
        dmins.append(float(info["Danger"].min_dist) if info["Danger"].val != 0 else np.nan)
        htimes.append(list(env.human_times)); gtimes.append(env.global_time)
        if done:
            break
    door = np.array([getattr(env, k, np.nan) for k in ("door_y_mid_min", "door_y_mid_max", "door_x_mid", "door_y_min",
                                                      "door_y_max", "door_width")], float)
    hp = env.humans[0].policy
    np.savez(os.path.join(OUT, f"rollout_{name}.npz"),
             human_policy=human_policy, sim=sim, case=case, H=H, unicycle=unicycle, starts_moving=starts_moving,
             segs=np.array(static, np.float64).reshape(-1, 4), h0=h0, r0=r0, human_times0=ht0, global_time0=gt0, robot_radius=robot.radius,
             robot_vpref=robot.v_pref,
             safety_space=getattr(hp, "safety_space", 0.0), policy_radius=getattr(hp, "radius", 0.0),
             is_bottleneck=bool(getattr(hp, "is_bottleneck", False)), door=door,
             time_limit=env.time_limit, time_step=env.time_step, circle_radius=env.circle_radius,
             H_states=np.array(H_states), R_states=np.array(R_states), actions=np.array(acts), reward=np.array(rew),
             flags=np.array(flags), dmin=np.array(dmins), human_times=np.array(htimes, float), global_time=np.array(gtimes))
    print(f"rollout {name}: steps={len(rew)} flags_or={np.bitwise_or.reduce(flags)} segs={len(static)}")


def gen_rollouts():
    rollout("orca_circle5", "orca", "circle_crossing", 5, 3, 60, sim__circle_radius=4.0, humans__radius=0.3)
    rollout("orca_circle10", "orca", "circle_crossing", 10, 0, 80, sim__circle_radius=4.0, humans__radius=0.3)
    rollout("orca_circle10_warm", "orca", "circle_crossing", 10, 1, 40, starts_moving=10, sim__circle_radius=4.0,
            humans__radius=0.3)
    rollout("orca_circle3_timeout", "orca", "circle_crossing", 3, 2, 30, sim__circle_radius=4.0, env__time_limit=5)
    rollout("orcaplus_hallway3", "orca_plus", "hallway", 3, 5, 60, starts_moving=10)
    rollout("orcaplus_static4", "orca_plus", "hallway_static", 4, 7, 60, sim__circle_radius=1.5, sim__rect_width=2.0)
    rollout("orcaplus_bottleneck4", "orca_plus", "hallway_bottleneck", 4, 9, 60, sim__circle_radius=1.5, sim__rect_width=2.0)
    rollout("orcaplus_squeeze3_uni", "orca_plus", "hallway_squeeze", 3, 4, 50, unicycle=True, sim__circle_radius=1.5,
            sim__rect_width=2.0)
    rollout("sfm_hallway5", "sfm", "hallway", 5, 11, 60, starts_moving=10)
    rollout("sfm_bottleneck5", "sfm", "hallway_bottleneck", 5, 13, 60, sim__circle_radius=1.5, sim__rect_width=2.0)
    rollout("sfm_static6_uni", "sfm", "hallway_static", 6, 17, 60, unicycle=True, sim__circle_radius=1.5,
            sim__rect_width=2.0)
    rollout("sfm_circle8", "sfm", "circle_crossing", 8, 21, 60, sim__circle_radius=4.0)
    zig = [(0.7, 0.35)] * 6 + [(0.0, 0.0)] * 3 + [(-0.7, 0.35)] * 12 + [(0.02, 0.0)] * 2 + [(0.7, 0.35)] * 12
    rollout("orcaplus_hallway6_zig", "orca_plus", "hallway", 6, 23, 70, starts_moving=10, pattern=zig)
    rollout("sfm_bottleneck6_zig", "sfm", "hallway_bottleneck", 6, 27, 70, pattern=zig, sim__circle_radius=1.5,
            sim__rect_width=2.0)
    rollout("orca_circle10_slow", "orca", "circle_crossing", 10, 31, 90, sim__circle_radius=4.0, humans__radius=0.3,
            robot__v_pref=0.4)


# ------------------------------------------------------------------ denoiser
def gen_jmid():
    import jmid_oracle as JO
    from sicnav_diffusion.JMID.MID.models import diffusion as D
    torch.manual_seed(0)
    out = {}
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for tag, ckpt, diffnet in (("jmid", "sim_gen_sicnav_p_midjp_cvg_epoch121.pt", "JointPredictionTransformerConcatLinear"),
                                   ("imid", "sim_gen_sicnav_p_mid_cvg_epoch169.pt", "TransformerConcatLinear")):
            dt, _ = ref_shims.load_jmid_reference(ckpt, diffnet)
            A, S, T = 2, 3, 8
            ctx = torch.linspace(-1, 1, A * 256).view(A, 256)
            x = torch.linspace(-1, 1, A * S * T * 2).view(A * S, T, 2)
            e = dt.net([x, ctx.repeat(S, 1)], beta=dt.var_sched.betas[[100] * (A * S)])
            smp, _ = dt.sample_sicnav_inference(T, ctx, S, bestof=False, sampling="ddim", step=20)
            out[f"{tag}_ckpt_kat_eps"] = e.numpy()
            out[f"{tag}_ckpt_kat_sample"] = smp.numpy()
            # injected x_T: replicate the loop of diffusion.py:507-531 around the reference net
            A, S = 3, 4
            ctx2 = torch.randn(A, 256, generator=g)
            xT = torch.randn(A * S, T, 2, generator=g)
            x_t = xT.clone()
            vs = dt.var_sched
            for t in range(100, 0, -5):
                ab, abn = vs.alpha_bars[t], vs.alpha_bars[t - 5]
                e_t = dt.net([x_t, ctx2.repeat(S, 1)], beta=vs.betas[[t] * (A * S)])
                x0 = (x_t - e_t * (1 - ab).sqrt()) / ab.sqrt()
                x_t = abn.sqrt() * x0 + (1 - abn).sqrt() * e_t
            out[f"{tag}_ckpt_ctx"] = ctx2.numpy(); out[f"{tag}_ckpt_xT"] = xT.numpy()
            out[f"{tag}_ckpt_sample20"] = x_t.reshape(S, A, T, 2).numpy()
            out[f"{tag}_alpha_bars"] = vs.alpha_bars.numpy(); out[f"{tag}_betas"] = vs.betas.numpy()

            # synthetic weights loaded INTO THE REFERENCE module (reproducible on the GPU box from the seed)
            w = JO.make_random_weights(seed=5)
            net = getattr(D, diffnet)(2, 256, 3, False)
            dt2 = D.DiffusionTraj(net, D.VarianceSchedule(num_steps=100, beta_T=5e-2, mode="linear"))
            sd = dt2.state_dict()
            for k, v in w.items():
                assert sd[k].shape == v.shape, k
                sd[k] = v
            dt2.load_state_dict(sd)
            dt2.eval()
            for (A, S, nm) in ((3, 4, "small"), (10, 20, "c4")):
                ctx3 = torch.randn(A, 256, generator=g)
                xT3 = torch.randn(A * S, T, 2, generator=g)
                e3 = dt2.net([xT3, ctx3.repeat(S, 1)], beta=dt2.var_sched.betas[[55] * (A * S)])
                x_t = xT3.clone()
                vs = dt2.var_sched
                steps = 20 if nm == "small" else 4
                stride = 100 // steps
                for t in range(100, 0, -stride):
                    ab, abn = vs.alpha_bars[t], vs.alpha_bars[t - stride]
                    e_t = dt2.net([x_t, ctx3.repeat(S, 1)], beta=vs.betas[[t] * (A * S)])
                    x0 = (x_t - e_t * (1 - ab).sqrt()) / ab.sqrt()
                    x_t = abn.sqrt() * x0 + (1 - abn).sqrt() * e_t
                out[f"{tag}_rand_{nm}_ctx"] = ctx3.numpy(); out[f"{tag}_rand_{nm}_xT"] = xT3.numpy()
                out[f"{tag}_rand_{nm}_eps55"] = e3.numpy()
                out[f"{tag}_rand_{nm}_sample"] = x_t.reshape(S, A, T, 2).numpy()
                out[f"{tag}_rand_{nm}_steps"] = np.array(steps)
    np.savez_compressed(os.path.join(OUT, "jmid_cases.npz"), rand_seed=5, **out)
    print("jmid:", sorted(out.keys()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["sfm", "geometry", "rollouts", "jmid"]
    with np.errstate(all="ignore"):
        if "sfm" in which:
            gen_sfm()
        if "geometry" in which:
            gen_geometry()
        if "rollouts" in which:
            gen_rollouts()
        if "jmid" in which:
            gen_jmid()


# ------------------------------------------------------------------ full predictor (a13-a17, a22-a25)
class _S:
    def __init__(self, x, y):
        self.position = (x, y)


def _pred_setup():
    import ref_predictor_shims as RPS
    EasyDict = RPS.install()
    from sicnav_diffusion.JMID import mid_sim_wrapper as W
    return EasyDict, W


def _pred_build(EasyDict, W, H, num_draw, num_ret, step):
    import yaml
    REF = ref_shims.REF
    cfg = configparser.RawConfigParser()
    cfg.read(REF_CFG)
    cfg.set("sim", "human_num", str(H))
    cfg.set("human_trajectory_forecaster", "num_samples", str(num_ret))
    y = yaml.safe_load(open(os.path.join(REF, "sicnav_diffusion/JMID/test_time_configs/mid_jp.yaml")))
    y.update(device="cpu", model_path=os.path.join(REF, y["model_path"]), num_samples=num_draw, step_size=step)
    return W.HumanTrajectoryForecasterSim(cfg, EasyDict(y))


def _pred_run(out, tag, f, H, rng, spread, synthetic, eps_at=()):
    """Feeds 8 frames of seeded quadratic motion, hooks the sampler to record ctx / x_T (and the raw sampled velocities), runs
    predict_ret_best.  eps_at: diffusion steps t at which the reference noise net is also evaluated on (x_T, ctx)."""
    import jmid_oracle as JO
    import predictor_oracle as PO
    if synthetic:
        ew = PO.make_random_encoder_weights(seed=9)
        md = f.mid_model.registrar.model_dict
        with torch.no_grad():
            for k, v in ew.items():
                mod, pname = k.rsplit("/", 1)
                dict(md[mod].named_parameters())[pname].copy_(v)
            dw = JO.make_random_weights(5)
            sd = f.mid_model.model.vel_predictor.state_dict()
            for k, v in dw.items():
                sd[k].copy_(v)
    p0 = rng.uniform(-spread, spread, (H, 2)); v0 = rng.uniform(-0.8, 0.8, (H, 2))
    rp = np.array([0.0, -spread * 0.6]); rv = np.array([0.1, 0.9])
    acc = rng.uniform(-0.3, 0.3, (H, 2))
    for i in range(H):
        f.prev_states[i].clear()
    f.prev_robot_states.clear()
    for k in range(8):          # more than 6 frames: the ring keeps the last 6
        t = 0.25 * k
        f.update_state_hists(_S(*(rp + rv * t)), [_S(*(p0[i] + v0[i] * t + 0.5 * acc[i] * t * t)) for i in range(H)], t)
    rec = {}
    vp = f.mid_model.model.vel_predictor
    orig_sample = vp.sample_sicnav_inference
    orig_randn = torch.randn

    def sample_hook(num_points, context, sample, bestof, **kw):
        rec["ctx"] = context.detach().clone()
        g = torch.Generator().manual_seed(4321)

        def fake_randn(*size, **k2):
            size = size[0] if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)) else size
            t_ = orig_randn(*size, generator=g)
            if "xT" not in rec:
                rec["xT"] = t_.clone()
            return t_
        torch.randn = fake_randn
        try:
            r = orig_sample(num_points, context, sample, bestof, **kw)
            rec["vel"] = r[0].detach().clone()
            return r
        finally:
            torch.randn = orig_randn
    vp.sample_sicnav_inference = sample_hook
    env_, poses, ids_in, ids_out, cv = f.convert_to_mid_state_env(f.prev_states, f.prev_robot_states)
    fc, lw = f.predict_ret_best()
    vp.sample_sicnav_inference = orig_sample
    out[f"{tag}_hist"] = np.array([s_ for s_ in f.prev_states], np.float64)          # [H,6,3]
    out[f"{tag}_robot_hist"] = np.array(f.prev_robot_states[-6:], np.float64)       # [6,3]
    out[f"{tag}_ids_in"] = np.array(ids_in, np.int64); out[f"{tag}_ids_out"] = np.array(ids_out, np.int64)
    out[f"{tag}_ctx"] = rec["ctx"].numpy(); out[f"{tag}_xT"] = rec["xT"].numpy()
    out[f"{tag}_forecasts"] = fc; out[f"{tag}_logw"] = lw
    out[f"{tag}_cfg"] = np.array([H, f.mid_model.num_samples, f.num_ret_samples, f.mid_model.config.step_size])
    if eps_at:
        S = f.mid_model.num_samples
        out[f"{tag}_vel"] = rec["vel"].numpy()                                      # [S,A,T,2] raw sampled velocities
        ctx_rep = rec["ctx"].repeat(S, 1)
        for t in eps_at:
            e = vp.net([rec["xT"], ctx_rep], beta=vp.var_sched.betas[[t] * ctx_rep.shape[0]])
            out[f"{tag}_eps{t}"] = e.numpy()
    print(f"predictor {tag}: H={H} in={list(ids_in)} out={list(ids_out)} ctx={tuple(rec['ctx'].shape)} fc={fc.shape}")


def gen_predictor():
    """Runs the reference's own HumanTrajectoryForecasterSim (mid_sim_wrapper.py) on CPU with (a) the shipped JMID
    checkpoint and (b) seeded synthetic encoder + denoiser weights written INTO the reference modules, recording the
    histories, the cluster split, the encoder context, the injected noise and predict_ret_best's outputs."""
    import predictor_oracle as PO
    EasyDict, W = _pred_setup()
    cwd = os.getcwd()
    os.chdir(ref_shims.REF)
    out = {}
    build = lambda *a: _pred_build(EasyDict, W, *a)
    run = lambda *a: _pred_run(out, *a)
    with torch.no_grad():
        rng = np.random.default_rng(2024)
        f = build(5, 20, 20, 20)
        md0 = f.mid_model.registrar.model_dict            # the shipped encoder weights (1 MB) so the ckpt cases pin encode()
        for k in PO.make_random_encoder_weights(seed=9):
            mod, pname = k.rsplit("/", 1)
            out["ckpt_enc:" + k] = dict(md0[mod].named_parameters())[pname].detach().numpy().copy()
        run("ckpt_h5", f, 5, rng, 1.5, False)
        run("ckpt_h5_sparse", f, 5, rng, 4.0, False)
        run("rand_h5", f, 5, rng, 1.5, True)
        run("rand_h5_sparse", f, 5, rng, 4.0, True)
        f = build(10, 20, 20, 4)
        run("rand_h10", f, 10, rng, 2.5, True)
        f = build(4, 20, 8, 5)            # num_ret < drawn: KDE top-k branch (get_most_likely_samples)
        run("rand_h4_kde", f, 4, rng, 1.2, True)
        run("ckpt_h4_kde", build(4, 20, 8, 5), 4, rng, 1.2, False)
    os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "predictor_cases.npz"), enc_seed=9, ddpm_seed=5, **out)


def gen_ckpt():
    """The SHIPPED JMID checkpoint (sim_gen_sicnav_p_midjp_cvg_epoch121.pt) for the GPU box:
      tests/golden/ckpt_jmid_epoch121.npz   its tensors ("ddpm/<state_dict key>", "enc/<module>/<param>"; the reference tree and
                                            its pickled nn.Modules do not travel), loadable by snb.jmid.load_checkpoint;
      tests/golden/ckpt_c4_cases.npz        outputs of the reference's own HumanTrajectoryForecasterSim / DiffusionTraj with that
                                            checkpoint at the C4 per-env shape (10 humans, 20 samples, 20 DDIM iterations, injected
                                            x_T): eps at t = 100 / 55 / 5, the sampled velocities, predict_ret_best."""
    EasyDict, W = _pred_setup()
    cwd = os.getcwd()
    os.chdir(ref_shims.REF)
    out = {}
    with torch.no_grad():
        rng = np.random.default_rng(777)
        f = _pred_build(EasyDict, W, 10, 20, 20, 20)
        wts = {}
        used_enc = ("PEDESTRIAN/node_history_encoder", "PEDESTRIAN->PEDESTRIAN/edge_encoder", "PEDESTRIAN->JRDB_ROBOT/edge_encoder",
                    "PEDESTRIAN/edge_influence_encoder")   # the modules get_latent touches (mgcvae.py:505-880); the rest is training-only
        for mod_name, mod in f.mid_model.registrar.model_dict.items():
            if mod_name in used_enc:
                for pname, t in mod.named_parameters():
                    wts[f"enc/{mod_name}/{pname}"] = t.detach().numpy().copy()
        for k, v in f.mid_model.model.vel_predictor.state_dict().items():
            if k.startswith("net.layer."):     # the template layer nn.TransformerEncoder deep-copied (diffusion.py:161-166); never run
                continue
            wts["ddpm/" + k] = v.detach().numpy().copy()
        np.savez(os.path.join(OUT, "ckpt_jmid_epoch121.npz"), **wts)
        _pred_run(out, "ckpt_h10_dense", f, 10, rng, 1.0, False, eps_at=(100, 55, 5))   # every human inside one cluster
        _pred_run(out, "ckpt_h10", f, 10, rng, 2.5, False, eps_at=(55,))
        f = _pred_build(EasyDict, W, 3, 100, 15, 2)      # the shipped simulation setting: 100 drawn, 2 DDIM iterations, 15 kept
        _pred_run(out, "ckpt_h3_shipped", f, 3, rng, 1.2, False, eps_at=(100,))
    os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "ckpt_c4_cases.npz"), **out)


def gen_ingest():
    """tests/golden/ingest_cases.npz: SICNavAcados.convert_to_mpc_state_vector (sicnav_acados.py:222-289) EXECUTED from the reference
    source.  The module cannot be imported (casadi / acados), so the method's own source text is cut out of the file with `ast` at
    generation time and exec'd against a stub `self` carrying the attributes it reads (mpc_env.nx_r / np_g / nX_hums / nx_hum /
    num_hums / num_MID_samples, human_pred_MID, human_pred_MID_joint, prev_rev)."""
    import ast
    import types
    src_path = os.path.join(ref_shims.REF, "sicnav_diffusion/policy/sicnav_acados.py")
    src = open(src_path).read()
    fn = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == "convert_to_mpc_state_vector":
            fn = node
    assert fn is not None
    ns = {"np": np}
    import textwrap
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    convert = ns["convert_to_mpc_state_vector"]
    rng = np.random.default_rng(99)
    out = {}
    for tag, joint, H, k in (("jmid_h10_k20", True, 10, 20), ("jmid_h3_k15", True, 3, 15), ("imid_h5_k8", False, 5, 8)):
        nx_hum = 6 if joint else 6 + k
        mpc_env = types.SimpleNamespace(nx_r=8, np_g=2, nx_hum=nx_hum, num_hums=H, num_MID_samples=k,
                                        nX_hums=nx_hum * H + (k if joint else 0))
        self_ = types.SimpleNamespace(mpc_env=mpc_env, human_pred_MID=True, human_pred_MID_joint=joint, prev_rev=False)
        robot = rng.uniform(-2, 2, 9)
        humans = rng.uniform(-3, 3, (H, 4)); humans[1, 2:] = 0.0            # a human at rest: theta = 0 branch (:1678)
        goals = rng.uniform(-4, 4, (H, 2))
        w = np.log(rng.dirichlet(np.ones(k), size=None if joint else H))
        rs = types.SimpleNamespace(px=robot[0], py=robot[1], theta=robot[2], lvel=robot[3], omega=robot[4], v_dot=robot[5],
                                   omega_dot=robot[6], gx=robot[7], gy=robot[8], velocity=(0.0, 0.0))
        hs = [types.SimpleNamespace(px=h[0], py=h[1], vx=h[2], vy=h[3], gx=g[0], gy=g[1]) for h, g in zip(humans, goals)]
        state = types.SimpleNamespace(self_state=rs, human_states=hs)
        val = convert(self_, state, 8, 2, mpc_env.nX_hums, w, get_numpy=True)
        out[f"{tag}_robot"] = robot; out[f"{tag}_humans"] = humans; out[f"{tag}_goals"] = goals; out[f"{tag}_weights"] = w
        out[f"{tag}_val"] = np.asarray(val).reshape(-1); out[f"{tag}_joint"] = np.array(int(joint))
    np.savez_compressed(os.path.join(OUT, "ingest_cases.npz"), **out)
    print("ingest:", sorted(out))


if "ingest" in sys.argv[1:]:
    gen_ingest()
if "predictor" in sys.argv[1:]:
    with np.errstate(all="ignore"):
        gen_predictor()
if "ckpt" in sys.argv[1:]:
    with np.errstate(all="ignore"):
        gen_ckpt()
