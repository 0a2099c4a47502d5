"""oracle/c1_flow.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

BASELINE configs[0] ("simple_test.py ..., 1 env on CPU: reference plumbing") as a restated loop: what simple_test.py:109-269 sets in
motion for ONE environment -- `policy_factory[name]()` per human, `policy.configure(config, 'humans')` inside try / except
(human_plus.py:12-16), then per step the object-at-a-time flow of CrowdSimPlus.step (crowd_sim_plus.py:1044-1055, 1193-1201):

    for human in humans:
        ob = [other.get_observable_state() for other in humans if other != human] + [robot.get_observable_state()]
        action = human.act(ob, static_obstacles)      # -> JointState(full_state, ob, static_obs) -> policy.predict(state)
        action = constrain_agent_action_exact(human, action)
    robot.step(constrained robot action); human.step(action) ...; human.set_g_xy()

with the POLICY OBJECTS INJECTED: the tests pass snb.policy.policy_factory (the drop-ins whose predict() is one B = 1 call of the
CUDA kernel through snb_policy_predict_host), and the loop must then reproduce a reference-generated golden episode.  The clamp and
the segment geometry come from the C oracle (pinned to the reference by tests/golden/geometry_cases.npz); kinematics and the door
goal are restated from agent_plus.py:175-214 / human_plus.py:19-52.

Why a restatement: the real CrowdSimPlus needs /root/reference (absent on the GPU box) and the drop-ins need a GPU (absent in the
build container), so "reference loop + swapped policy_factory entries" cannot execute in either place; this file is that loop.
"""
import ctypes as C
import math

import numpy as np

import oracle_lib as ol
import scenario_oracle as SO


def _constrain(px, py, theta, radius, dt, unicycle, a0, a1, segs):
    """CrowdSimPlus.constrain_agent_action_exact (crowd_sim_plus.py:869-989) via the C oracle."""
    if len(segs) == 0:
        return a0, a1
    pose = np.array([px, py, theta], np.float64)
    act = np.array([a0, a1], np.float64)
    out = np.zeros(2, np.float64)
    flat = np.ascontiguousarray(np.asarray(segs, np.float64).reshape(-1))
    ol.lib().orc_constrain_action(ol.dptr(pose), float(radius), float(dt), int(unicycle), ol.dptr(act), len(segs), ol.dptr(flat), ol.dptr(out))
    return float(out[0]), float(out[1])


def run_episode(g, policy_factory, state_classes, env_config):
    """g: a golden rollout (tests/golden/rollout_*.npz, loaded by golden_util.load_rollout).  policy_factory: name -> Policy class,
    state_classes = (FullState, ObservableState, JointState).  Yields after every step (human_states [H,7] = px,py,vx,vy,theta,gx,gy,
    robot_state [7] = px,py,vx,vy,theta,gx,gy)."""
    FullState, ObservableState, JointState = state_classes
    H = int(g["H"])
    dt = float(g["time_step"])
    segs = [[(s[0], s[1]), (s[2], s[3])] for s in g["segs"]]
    segs4 = np.asarray(g["segs"], np.float64).reshape(-1, 4)
    door = None
    if g["sim"] in SO.DOOR_RULES and len(segs):
        d = g["door"]
        door = dict(door_y_mid_min=d[0], door_y_mid_max=d[1], door_x_mid=d[2], door_y_min=d[3], door_y_max=d[4], door_width=d[5])
    # Human.__init__ (human_plus.py:6-17): one policy object per human, configure swallowed on failure (quirk q6)
    policies = []
    for _ in range(H):
        p = policy_factory[g["human_policy"]]()
        try:
            p.configure(env_config, 'humans')
        except Exception:
            pass
        p.time_step = dt                                            # CrowdSimPlus.reset, crowd_sim_plus.py:686-688
        if g["human_policy"] == "sfm":
            p.is_bottleneck = bool(g["is_bottleneck"])               # crowd_sim_plus.py:448-449
        policies.append(p)
    hs = [dict(px=r[0], py=r[1], vx=r[2], vy=r[3], theta=r[4], gx=r[5], gy=r[6], fgx=r[7], fgy=r[8], v_pref=r[9], radius=r[10])
          for r in np.asarray(g["h0"], np.float64)]
    r0 = np.asarray(g["r0"], np.float64)
    rb = dict(px=r0[0], py=r0[1], vx=r0[2], vy=r0[3], theta=r0[4], gx=r0[5], gy=r0[6], radius=float(g["robot_radius"]))
    unicycle = bool(g["unicycle"])
    for k in range(len(g["actions"])):
        acts = []
        for i, h in enumerate(hs):
            ob = [ObservableState(o["px"], o["py"], o["vx"], o["vy"], o["radius"]) for j, o in enumerate(hs) if j != i]
            ob += [ObservableState(rb["px"], rb["py"], rb["vx"], rb["vy"], rb["radius"])]                 # robot last (:1047-1049)
            me = FullState(h["px"], h["py"], h["vx"], h["vy"], h["radius"], h["gx"], h["gy"], h["v_pref"], h["theta"])
            a = policies[i].predict(JointState(me, ob, segs))                                             # human.act -> policy.predict
            acts.append(_constrain(h["px"], h["py"], h["theta"], h["radius"], dt, False, float(a.vx), float(a.vy), segs4))
        ra = _constrain(rb["px"], rb["py"], rb["theta"], rb["radius"], dt, unicycle, float(g["actions"][k][0]), float(g["actions"][k][1]), segs4)
        # Agent.compute_position + Agent.step (agent_plus.py:175-185, 199-214)
        if unicycle:
            th = rb["theta"] + ra[1]
            rb["px"], rb["py"] = rb["px"] + np.cos(th) * ra[0] * dt, rb["py"] + np.sin(th) * ra[0] * dt
            un = th % (2 * np.pi)
            rb["theta"] = un - 2 * np.pi if un > np.pi else un
            rb["vx"], rb["vy"] = ra[0] * np.cos(rb["theta"]), ra[0] * np.sin(rb["theta"])
        else:
            rb["px"], rb["py"] = rb["px"] + ra[0] * dt, rb["py"] + ra[1] * dt
            rb["vx"], rb["vy"] = ra
            rb["theta"] = np.arctan2(ra[1], ra[0])
        for h, (vx, vy) in zip(hs, acts):
            h["px"] += vx * dt; h["py"] += vy * dt
            h["vx"], h["vy"] = vx, vy
            h["theta"] = np.arctan2(vy, vx)
            h["gx"], h["gy"] = SO.door_goal(g["sim"], door, len(segs), h["px"], h["py"], h["fgx"], h["fgy"])   # Human.step -> set_g_xy
        yield (np.array([[h[n] for n in ("px", "py", "vx", "vy", "theta", "gx", "gy")] for h in hs]),
               np.array([rb[n] for n in ("px", "py", "vx", "vy", "theta", "gx", "gy")]))
