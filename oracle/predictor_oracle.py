"""oracle/predictor_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + torch tensor algebra) of the JMID predictor around the denoiser, i.e. everything
HumanTrajectoryForecasterSim.predict_ret_best does (paths relative to the reference root):
  * history -> per-node state (finite differences), 3 m clustering, constant-velocity fall-back
      sicnav_diffusion/JMID/mid_sim_wrapper.py:244-437, MID/environment/data_utils.py:24-37
  * scene graph + edge scaling, standardisation, neighbour batches
      MID/environment/scene_graph.py:111-225, 280-313; MID/dataset/preprocessing.py:428-620
  * Trajectron context encoder (history LSTM, per-edge-type LSTM, additive attention)
      MID/models/encoders/mgcvae.py:505-880, components/additive_attention.py:6-47
  * KDE top-k (get_most_likely_samples, mid_sim_wrapper.py:14-169), scatter + current pose (:482-509)
  * the MPC ingest of sicnav_diffusion/policy/sicnav_acados.py:1645-1667
Pinned against the reference's own stack run on CPU in the build container (tests/golden/predictor_cases.npz, made by
oracle/gen_golden.py predictor) with the shipped checkpoint and with seeded synthetic weights.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import math

import numpy as np
import torch

import jmid_oracle as JO

ENC_KEYS = {
    "PEDESTRIAN/node_history_encoder": 6,
    "PEDESTRIAN->PEDESTRIAN/edge_encoder": 12,
    "PEDESTRIAN->JRDB_ROBOT/edge_encoder": 12,
}
STD = np.array([3.0, 3.0, 2.0, 2.0, 1.0, 1.0])   # position std := the attention radius, 3.0 m as shipped (preprocessing.py:477-478)


def make_random_encoder_weights(seed=9):
    """Seeded synthetic encoder weights with the names of `checkpoint["encoder"]`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    b = 1.0 / math.sqrt(128)
    for mod, din in ENC_KEYS.items():
        w[f"{mod}/weight_ih_l0"] = rng.uniform(-b, b, (512, din))
        w[f"{mod}/weight_hh_l0"] = rng.uniform(-b, b, (512, 128))
        w[f"{mod}/bias_ih_l0"] = rng.uniform(-b, b, (512,))
        w[f"{mod}/bias_hh_l0"] = rng.uniform(-b, b, (512,))
    for n, s in (("w1.weight", (128, 128)), ("w2.weight", (128, 128)), ("v.weight", (1, 128))):
        w[f"PEDESTRIAN/edge_influence_encoder/{n}"] = rng.uniform(-b, b, s)
    return {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in w.items()}


def derivative_of(x, dt):
    """data_utils.derivative_of: first difference duplicated at the front, divided by dt."""
    return np.ediff1d(x, to_begin=(x[1] - x[0])) / dt


def node_states(pos, dt):
    """pos [Th,2] -> [Th,6] = x,y,vx,vy,ax,ay (mid_sim_wrapper.py:376-391)."""
    x, y = pos[:, 0], pos[:, 1]
    vx, vy = derivative_of(x, dt), derivative_of(y, dt)
    ax, ay = derivative_of(vx, dt), derivative_of(vy, dt)
    return np.stack([x, y, vx, vy, ax, ay], 1)


def cluster_split(hist, robot_hist, radius=3.0):
    """mid_sim_wrapper.py:322-355.  hist [H,Th,3], robot_hist [Th,3] -> (in_cluster bool [H+1], order) with index 0 = robot
    (track id -1) and index 1+i = human i, exactly the sorted-by-track-id order of the reference."""
    positions = np.concatenate([robot_hist[-1:, :2], hist[:, -1, :2]], 0)
    sq = np.square(positions[:, None] - positions[None, :])
    dists = np.sqrt(np.sum(sq, axis=2))
    mask = dists < radius
    cluster_means = (mask @ positions) / mask.sum(axis=1, keepdims=True)
    robot_dist = np.linalg.norm(cluster_means - positions[0], axis=1)
    chosen = int(np.argmin(robot_dist[1:]) + 1)
    return mask[chosen]


def edge_scaling(pos3, is_ped, radius=3.0):
    """TemporalSceneGraph.create_from_temp_scene_dict + calculate_edge_scaling at the current frame for the shipped
    filters add=[.25,.5,.75,1], remove=[1,0] (scene_graph.py:111-225).  pos3 [3,n,2] = frames t-2..t."""
    n = pos3.shape[1]
    adj = np.zeros((3, n, n))
    for k in range(3):
        d = np.sqrt(((pos3[k][:, None] - pos3[k][None, :]) ** 2).sum(-1))
        a = (d <= radius).astype(np.float64)
        np.fill_diagonal(a, 0)
        adj[k] = a
    new_edges = np.minimum(0.25 * adj[2] + 0.5 * adj[1] + 0.75 * adj[0], 1.0)
    new_edges[adj[2] == 0] = 0
    return new_edges        # removal filter [1, 0] leaves the current frame unchanged


def encoder_inputs(hist, robot_hist, dt=0.25, radius=3.0, pos_std=None):
    """Everything up to the encoder: returns dict(in_cluster [H] bool, ped_ids (in-cluster humans, ascending),
    x_st [A,6,6], nb_ped [A,6,6], nb_rob [A,6,6], edge_mask [A], p0 [A,2], cv {h: [T,2]} filled by the caller)."""
    H, Th, _ = hist.shape
    std = STD.copy()
    std[0:2] = radius if pos_std is None else pos_std        # std[0:2] = env.attention_radius[...] (preprocessing.py:478, 540)
    inc = cluster_split(hist, robot_hist, radius)
    nodes = [i for i in range(H + 1) if inc[i]]             # 0 = robot
    states = {}
    for i in nodes:
        p = robot_hist[-Th:, :2] if i == 0 else hist[i - 1, :, :2]
        states[i] = node_states(p, dt)
    pos3 = np.stack([np.stack([states[i][Th - 3 + k, :2] for i in nodes], 0) for k in range(3)], 0)
    es = edge_scaling(pos3, None, radius)
    ped_nodes = [i for i in nodes if i != 0]
    x_st, nb_ped, nb_rob, emask, p0 = [], [], [], [], []
    for a, i in enumerate(nodes):
        if i == 0:
            continue
        x = states[i]
        rel = np.zeros(6); rel[0:2] = x[-1, 0:2]
        x_st.append((x - rel) / std)
        conn = es[a] > 1e-2                                   # SceneGraph.get_connection_mask
        sp, sr = np.zeros((Th, 6)), np.zeros((Th, 6))
        for b_, j in enumerate(nodes):
            if not conn[b_]:
                continue
            nst = (states[j] - x[-1][None, :]) / std         # relative to the node's CURRENT full state (:541-551)
            if j == 0:
                sr += nst
            else:
                sp += nst
        nb_ped.append(sp); nb_rob.append(sr)
        emask.append(min(float(es[a][conn].sum()), 1.0))     # clamp(sum over ALL neighbours, 1) for every edge type (q5)
        p0.append(x[-1, 0:2])
    return dict(in_cluster=inc[1:], ped_ids=[i - 1 for i in ped_nodes], x_st=np.array(x_st), nb_ped=np.array(nb_ped),
                nb_rob=np.array(nb_rob), edge_mask=np.array(emask), p0=np.array(p0))


def _lstm_last(w, mod, seq):
    """nn.LSTM(batch_first) over [A,Th,in] from zero state, last output (model_utils.py:77-105 with full histories)."""
    wih, whh = w[f"{mod}/weight_ih_l0"], w[f"{mod}/weight_hh_l0"]
    b = w[f"{mod}/bias_ih_l0"] + w[f"{mod}/bias_hh_l0"]
    A = seq.shape[0]
    h = torch.zeros(A, 128); c = torch.zeros(A, 128)
    for t in range(seq.shape[1]):
        g = seq[:, t] @ wih.T + h @ whh.T + b
        i, f, gg, o = g[:, :128], g[:, 128:256], g[:, 256:384], g[:, 384:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


def encode(w, inp):
    """MultimodalGenerativeCVAE.obtain_encoded_tensors in PREDICT mode (mgcvae.py:505-681): ctx [A,256] fp32."""
    x_st = torch.tensor(inp["x_st"], dtype=torch.float32)
    hist = _lstm_last(w, "PEDESTRIAN/node_history_encoder", x_st)
    m = torch.tensor(inp["edge_mask"], dtype=torch.float32).view(-1, 1)
    edges = []
    for mod, key in (("PEDESTRIAN->PEDESTRIAN/edge_encoder", "nb_ped"), ("PEDESTRIAN->JRDB_ROBOT/edge_encoder", "nb_rob")):
        joint = torch.cat([torch.tensor(inp[key], dtype=torch.float32), x_st], dim=-1)
        edges.append(_lstm_last(w, mod, joint) * m)
    a = "PEDESTRIAN/edge_influence_encoder/"
    scores = torch.cat([torch.tanh(e @ w[a + "w1.weight"].T + hist @ w[a + "w2.weight"].T) @ w[a + "v.weight"].T for e in edges], dim=1)
    p = torch.softmax(scores, dim=1)
    combined = p[:, 0:1] * edges[0] + p[:, 1:2] * edges[1]
    return torch.cat([combined, hist], dim=1)


def kde_totals(forecasts):
    """Per-sample total log-likelihood of get_most_likely_samples (mid_sim_wrapper.py:14-150), forecasts [S,A,T,2] -> [S].
    NB with the shipped bandwidths (0.01 .. 0.1, applied on top of the covariance whitening) the kernel is so narrow that every
    sample only sees itself unless two samples nearly coincide: all totals are then EXACTLY T*log(1/S) and the reference's
    top-k is decided by the tie order of torch.argsort."""
    S, A, T, _ = forecasts.shape
    preds = forecasts.permute(2, 0, 1, 3).reshape(T, S, A * 2)
    bandwidth = torch.exp(torch.linspace(np.log(0.01), np.log(0.1), steps=T))
    n, d = torch.tensor(float(S)), 2 * A
    diff = preds - preds.mean(dim=1, keepdim=True)
    cov = torch.bmm(diff.transpose(1, 2), diff) / (n - 1)
    sci = bandwidth[:, None, None] ** -2 * cov + torch.eye(d).expand_as(cov) * 1e-6
    L = torch.linalg.cholesky_ex(torch.inverse(sci))[0]
    diffs = preds.unsqueeze(2) - preds.unsqueeze(1)
    diffs = torch.matmul(diffs, torch.linalg.inv(L).unsqueeze(1)) / bandwidth[:, None, None, None]
    log_exp = -0.5 * torch.norm(diffs, p=2, dim=-1) ** 2
    log_det = 2 * torch.sum(torch.log(torch.diagonal(L, dim1=-2, dim2=-1)), dim=-1)
    Z = 0.5 * d * torch.log(torch.tensor(2 * np.pi)) + 0.5 * log_det.unsqueeze(-1) + torch.log(n)
    ll = torch.logsumexp(log_exp - Z.unsqueeze(-1), dim=-1)
    ll = ll - torch.logsumexp(ll, dim=1, keepdim=True)
    return ll.sum(0)


def most_likely_samples(forecasts, k):
    """get_most_likely_samples, joint branch (mid_sim_wrapper.py:14-169; quirk q4: always joint).
    forecasts [S,A,T,2] torch -> ([A,k,T,2], logw [A,k])."""
    S, A, T, _ = forecasts.shape
    tot = kde_totals(forecasts)
    top = torch.argsort(tot)[-k:]
    lw = tot[top] - torch.logsumexp(tot[top], dim=-1, keepdim=True)
    return forecasts[top].permute(1, 0, 2, 3), lw.unsqueeze(0).expand(A, k)


def predict_ret_best(enc_w, ddpm_w, hist, robot_hist, x_T, num_draw, num_ret, step, dt=0.25, horizon=8, joint=True, radius=3.0,
                     pos_std=None):
    """HumanTrajectoryForecasterSim.predict_ret_best (mid_sim_wrapper.py:482-509) with injected noise x_T [S*A,T,2].
    Returns (forecasts [H,k,T+1,2] float64, logw [H,k] float64, ctx [A,256])."""
    H = hist.shape[0]
    inp = encoder_inputs(hist, robot_hist, dt, radius, pos_std)
    ctx = encode(enc_w, inp)
    vel = JO.sample(ddpm_w, ctx, x_T, step=step, joint=joint)                       # [S,A,T,2]
    pos = JO.integrate(vel, torch.tensor(inp["p0"], dtype=torch.float32), dt)       # ascending human id == sort by node id
    if num_ret < num_draw:
        fc_in, lw_in = most_likely_samples(pos, num_ret)
        fc_in, lw_in = fc_in.numpy(), lw_in.numpy()
    else:
        fc_in = pos.permute(1, 0, 2, 3).numpy()
        lw_in = np.log(np.ones((pos.shape[1], num_draw), np.float64) / num_draw)
    forecasts = np.zeros((H, num_ret, horizon, 2), np.float64)
    logw = np.zeros((H, num_ret), np.float64)
    forecasts[inp["ped_ids"]] = fc_in
    logw[inp["ped_ids"]] = lw_in
    for h in range(H):
        if inp["in_cluster"][h]:
            continue
        st = node_states(hist[h, :, :2], dt)                                         # constant-velocity fall-back (:413-429)
        fc = np.zeros((horizon, 2))
        fc[:, 0] = st[-1, 0] + np.cumsum(np.tile(st[-1, 2] * dt, horizon))
        fc[:, 1] = st[-1, 1] + np.cumsum(np.tile(st[-1, 3] * dt, horizon))
        forecasts[h] = fc[None]
        logw[h] = lw_in[0]
    cur = np.repeat(hist[:, -1, None, None, :2], num_ret, axis=1)                  # add_current_pose_to_forecasts (:444-454)
    return np.concatenate([cur, forecasts], axis=2), logw, ctx


def mpc_ingest(forecasts, logw, dt=0.25, horiz=4, joint=True):
    """SICNavAcados.predict ingest (sicnav_acados.py:1645-1667): returns (forecasts_reshaped [horiz+1, H*k, 2], weights,
    goals [H,2], v_pref [H])."""
    fc = forecasts[:, :, 1:, :]
    weights = logw[0, :] if joint else logw
    H, k, T, _ = fc.shape
    resh = np.transpose(fc, (2, 0, 1, 3)).reshape(T, H * k, 2)[:horiz + 1]
    goals = np.array([[np.mean(fc[h, :, 0, 0]), np.mean(fc[h, :, 0, 1])] for h in range(H)])    # per human, as the reference loops
    v = np.linalg.norm(np.diff(fc, axis=2), axis=-1) / dt                           # on the forecasts WITHOUT the current pose (:1667)
    return resh, weights, goals, v.max(axis=(1, 2))


def mpc_state_vector(robot, humans, goals, weights, joint=True):
    """SICNavAcados.convert_to_mpc_state_vector (sicnav_diffusion/policy/sicnav_acados.py:222-289) on the joint state that
    SICNavAcados.predict builds (:1655-1681).  robot = (px, py, theta, lvel, omega, v_dot, omega_dot, gx, gy); humans [H,4] =
    px, py, vx, vy; goals [H,2]; weights [k] (JMID) or [H,k] (iMID).  nx_r = 8, np_g = 2, nx_hum = 6 (+ k for iMID)
    (utils/mpc_utils/mpc_env_new.py:72-104).  Returns (val [nx], human_theta [H])."""
    H = humans.shape[0]
    k = weights.shape[-1]
    nx_hum = 6 if joint else 6 + k
    val = np.zeros(8 + 2 + nx_hum * H + (k if joint else 0))
    val[0], val[1] = robot[0], robot[1]
    val[2], val[3] = np.sin(robot[2]), np.cos(robot[2])
    val[4:8] = robot[3:7]
    val[8], val[9] = robot[7], robot[8]
    off = 10
    for i in range(H):
        val[off + i * nx_hum: off + i * nx_hum + 4] = humans[i]
        val[off + i * nx_hum + 4: off + i * nx_hum + 6] = goals[i]
        if not joint:
            val[off + i * nx_hum + 6: off + (i + 1) * nx_hum] = weights[i, :]
    if joint:
        val[-k:] = weights[:]
    theta = np.array([np.arctan2(h[3], h[2]) if (h[2] != 0 or h[3] != 0) else 0.0 for h in humans])      # :1678
    return val, theta


def stage_params(resh, horiz, prefix=None, static_obs=None):
    """The parameter vector set on every solver stage (sicnav_acados.py:1389-1413): np.hstack([x_ref_k, u_ref_k, Q, R, Q_T (the
    `prefix`, MPC-side), X_t[:,0], X_t[:,1], X_t+1[:,0], X_t+1[:,1] (, static obstacles)]); stage `horiz` reuses t = horiz - 1.
    resh = forecasts_reshaped [horiz+1, H*k, 2].  Returns [horiz+1, n_prefix + 4*H*k + n_static]."""
    rows = []
    for idx in range(horiz + 1):
        t = idx if idx < horiz else horiz - 1
        parts = [] if prefix is None else [prefix[idx]]
        parts += [resh[t][:, 0], resh[t][:, 1], resh[t + 1][:, 0], resh[t + 1][:, 1]]
        if static_obs is not None:
            parts.append(np.asarray(static_obs).reshape(-1))
        rows.append(np.hstack(parts))
    return np.stack(rows)


def bootstrap_history(states, num_hist_frames=6):
    """reset_scenario_values (sicnav_acados.py:1163-1182): a fresh forecaster is fed env.states[-Th-1:-1] -- the newest logged state
    is NOT used -- stamped (global_time_step - Th ... global_time_step - 1) * dt.  states: list of [H+1,2] position arrays (robot
    last), oldest first.  Returns (hist [H,Th,2], robot_hist [Th,2])."""
    if len(states) < num_hist_frames + 1:
        raise IndexError("not enough logged states (starts_moving too small)")
    sel = np.stack(states[-num_hist_frames - 1:-1])             # [Th, H+1, 2]
    return np.ascontiguousarray(sel[:, :-1].transpose(1, 0, 2)), np.ascontiguousarray(sel[:, -1])
