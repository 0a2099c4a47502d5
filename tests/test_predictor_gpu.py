"""GPU parity tests of the JMID predictor around the denoiser (run with -m gpu), through the C ABI (ctypes).

Checkers: oracle/predictor_oracle.py (pinned to the reference stack, tests/test_predictor_oracle_golden.py) and
tests/golden/predictor_cases.npz (outputs of the reference's HumanTrajectoryForecasterSim itself).

Tolerances:
  * cluster split, slot order, in-cluster flags, constant-velocity rows, current pose, uniform log-weights, MPC ingest:
    exact (fp64 / integer work, no FMA contraction)
  * encoder context (fp32 LSTMs, different summation order than torch): <= 2e-5 absolute (|ctx| <= 1)
  * forecasts of in-cluster humans (bf16 tensor-core denoiser): <= 2e-2 m
  * KDE top-k (fp32 like the reference, Gauss-Jordan instead of LU): a valid top-k set w.r.t. the oracle's totals (2e-2),
    log-weights <= 2e-2; exactly tied totals (the common case, see PO.kde_totals) -> uniform weights, index tie order
"""
import configparser

import numpy as np
import pytest

import jmid_oracle as JO
import predictor_oracle as PO
from golden_util import GOLDEN

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

G = np.load(f"{GOLDEN}/predictor_cases.npz")
CASES = ["ckpt_h5", "ckpt_h5_sparse", "rand_h5", "rand_h5_sparse", "rand_h10", "rand_h4_kde", "ckpt_h4_kde"]
_cache = {}


def enc_weights(tag):
    if tag.startswith("ckpt"):
        return {k[9:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("ckpt_enc:")}
    return PO.make_random_encoder_weights(int(G["enc_seed"]))


def forecaster(tag_kind, B, H, S=20, k=None, step=20):
    from snb.jmid.forecaster import ForecasterBatch
    key = (tag_kind, B, H, S, k, step)
    if key not in _cache:
        _cache.clear()          # one live predictor at a time (activation buffers are sized for the chunk)
        _cache[key] = ForecasterBatch(enc_weights(tag_kind), JO.make_random_weights(int(G["ddpm_seed"])), max_envs=B, H=H,
                                      num_samples=S, num_ret=k, step_size=step)
    return _cache[key]


def random_histories(B, H, seed, spread=3.0):
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(-spread, spread, (B, H + 1, 1, 2))
    v0 = rng.uniform(-0.9, 0.9, (B, H + 1, 1, 2))
    acc = rng.uniform(-0.4, 0.4, (B, H + 1, 1, 2))
    t = (0.25 * np.arange(6)).reshape(1, 1, 6, 1)
    pos = p0 + v0 * t + 0.5 * acc * t * t
    return np.ascontiguousarray(pos[:, 1:]), np.ascontiguousarray(pos[:, 0])


@pytest.mark.parametrize("tag", CASES)
def test_encoder_matches_reference_golden(tag):
    hist, rh = G[tag + "_hist"][..., :2], G[tag + "_robot_hist"][..., :2]
    H = hist.shape[0]
    f = forecaster(tag[:4], 1, H)
    f.set_history(torch.from_numpy(np.ascontiguousarray(hist[None])).cuda(), torch.from_numpy(np.ascontiguousarray(rh[None])).cuda())
    ctx, n_in, ped, inc = (t.cpu().numpy() for t in f.encode(1))
    ids_in = list(G[tag + "_ids_in"])
    assert n_in[0] == len(ids_in) and list(ped[0, :len(ids_in)]) == ids_in and all(ped[0, len(ids_in):] == -1)
    assert [h for h in range(H) if not inc[0, h]] == list(G[tag + "_ids_out"])
    assert np.max(np.abs(ctx[0, :len(ids_in)] - G[tag + "_ctx"])) <= 2e-5


def test_encoder_batched_random_scenes_match_oracle():
    B, H = 96, 10
    hist, rh = random_histories(B, H, seed=5)
    f = forecaster("rand", B, H)
    f.set_history(torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    ctx, n_in, ped, inc = (t.cpu().numpy() for t in f.encode(B))
    w = enc_weights("rand")
    sizes = set()
    for b in range(B):
        h3 = np.concatenate([hist[b], np.zeros((H, 6, 1))], -1); r3 = np.concatenate([rh[b], np.zeros((6, 1))], -1)
        inp = PO.encoder_inputs(h3, r3)
        A = len(inp["ped_ids"])
        sizes.add(A)
        assert n_in[b] == A and list(ped[b, :A]) == inp["ped_ids"], b
        assert np.array_equal(inc[b].astype(bool), inp["in_cluster"]), b
        ref = PO.encode(w, inp).numpy()
        assert np.max(np.abs(ctx[b, :A] - ref)) <= 2e-5, b
    assert len(sizes) >= 4      # the scenes really exercise different cluster sizes


def _noise_from_xT(xT, A, S, H, T=8):
    """reference x_T [S*A,T,2] (row s*A+a) -> the library's [1,S,H,T,2] slot layout"""
    nz = np.zeros((1, S, H, T, 2), np.float32)
    nz[0, :, :A] = xT.reshape(S, A, T, 2)
    return nz


@pytest.mark.parametrize("tag", ["rand_h5", "rand_h5_sparse", "rand_h10"])
def test_predict_ret_best_matches_reference_golden(tag):
    hist, rh = G[tag + "_hist"][..., :2], G[tag + "_robot_hist"][..., :2]
    H, n_draw, n_ret, step = (int(v) for v in G[tag + "_cfg"])
    f = forecaster("rand", 1, H, S=n_draw, k=n_ret, step=step)
    ids_in, ids_out = list(G[tag + "_ids_in"]), list(G[tag + "_ids_out"])
    nz = _noise_from_xT(G[tag + "_xT"], len(ids_in), n_draw, H)
    fc, lw = f.predict_host(hist[None], rh[None], nz)
    ref_fc, ref_lw = G[tag + "_forecasts"], G[tag + "_logw"]
    assert fc.shape == (1,) + ref_fc.shape
    assert np.array_equal(lw[0], ref_lw)                                          # log(1/S) everywhere
    assert np.array_equal(fc[0][:, :, 0], ref_fc[:, :, 0])                        # current pose
    assert np.array_equal(fc[0][ids_out], ref_fc[ids_out])                        # constant-velocity rows, bit-exact
    err = np.max(np.abs(fc[0][ids_in] - ref_fc[ids_in]))
    assert err <= 2e-2, err


def test_predict_batched_groups_equal_single_environment_runs():
    """Environments are grouped by cluster size and denoised group by group: every environment must come out exactly as when it
    is predicted alone."""
    B, H, S = 24, 6, 5
    hist, rh = random_histories(B, H, seed=11, spread=2.5)
    f = forecaster("rand", B, H, S=S, step=4)
    noise = torch.randn(B, S, H, 8, 2, generator=torch.Generator().manual_seed(1)).cuda()
    f.set_history(torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    fc, lw = f.predict(B, noise=noise)
    _, n_in, _, _ = f.encode(B)
    assert len(set(n_in.cpu().tolist())) >= 3
    fc, lw = fc.cpu().numpy(), lw.cpu().numpy()
    w_enc, w_dd = enc_weights("rand"), JO.make_random_weights(int(G["ddpm_seed"]))
    for b in (0, 7, 13, 23):
        f.set_history(torch.from_numpy(hist[b:b + 1].copy()).cuda(), torch.from_numpy(rh[b:b + 1].copy()).cuda())
        fc1, lw1 = f.predict(1, noise=noise[b:b + 1].contiguous())
        assert np.array_equal(fc1.cpu().numpy()[0], fc[b]) and np.array_equal(lw1.cpu().numpy()[0], lw[b]), b
        # and against the oracle run of that environment
        h3 = np.concatenate([hist[b], np.zeros((H, 6, 1))], -1); r3 = np.concatenate([rh[b], np.zeros((6, 1))], -1)
        A = int(n_in[b])
        xT = noise[b, :, :A].reshape(S * A, 8, 2).cpu()
        with torch.no_grad():
            ref_fc, ref_lw, _ = PO.predict_ret_best(w_enc, w_dd, h3, r3, xT, S, S, 4)
        assert np.max(np.abs(fc[b] - ref_fc)) <= 2e-2, b
        assert np.array_equal(lw[b], ref_lw)


def _clustered_samples(seed, S, A, scale, T=8):
    """groups of nearly coincident samples: the only regime in which the reference's KDE (bandwidth 0.01 .. 0.1 on top of the
    covariance whitening) separates the samples; otherwise every total is exactly T*log(1/S) (see PO.kde_totals)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(8, A, T, 2, generator=g)
    grp = torch.tensor([0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7])[torch.randperm(20, generator=g)][:S]
    return (base[grp] + torch.randn(S, A, T, 2, generator=g) * scale).contiguous()


@pytest.mark.parametrize("A,scale", [(1, 1e-4), (2, 1e-4), (2, 1e-3), (3, 3e-4)])
def test_kde_topk_kernel_matches_oracle(A, scale):
    from snb.jmid.forecaster import kde_topk
    B, S, k = 5, 20, 8
    pos = torch.stack([_clustered_samples(100 * A + b, S, A, scale) for b in range(B)])
    sel, lw = kde_topk(pos.cuda(), k)
    sel, lw = sel.cpu().numpy(), lw.cpu().numpy()
    for b in range(B):
        tot = PO.kde_totals(pos[b]).numpy().astype(np.float64)
        assert np.isfinite(tot).all()
        order = np.argsort(tot)
        kth = tot[order[-k]]
        assert len(set(sel[b])) == k
        assert all(tot[i] >= kth - 2e-2 for i in sel[b]), (b, sel[b], order[-k:])            # a valid top-k set
        assert all(tot[sel[b][j]] <= tot[sel[b][j + 1]] + 2e-2 for j in range(k - 1))       # ascending, like argsort()[-k:]
        ref = tot[sel[b]] - np.log(np.exp(tot[sel[b]] - tot[sel[b]].max()).sum()) - tot[sel[b]].max()
        assert np.max(np.abs(lw[b] - ref)) <= 2e-2, (b, lw[b], ref)
        assert abs(np.exp(lw[b]).sum() - 1.0) <= 1e-5


def test_kde_degenerate_totals_give_uniform_weights_and_index_order():
    """Well separated samples: all totals tie exactly (reference behaviour), weights are uniform log(1/k); the library breaks the
    tie by sample index (torch.argsort's tie order is unspecified)."""
    from snb.jmid.forecaster import kde_topk
    pos = torch.randn(3, 20, 2, 8, 2, generator=torch.Generator().manual_seed(5))
    sel, lw = kde_topk(pos.cuda(), 8)
    assert np.array_equal(sel.cpu().numpy(), np.tile(np.arange(12, 20), (3, 1)))
    assert np.max(np.abs(lw.cpu().numpy() - np.log(1.0 / 8))) <= 1e-6
    tot = PO.kde_totals(pos[0])
    assert torch.all(tot == tot[0])


def test_predict_with_fewer_returned_samples_selects_drawn_samples():
    """num_ret < drawn through the whole pipeline: the returned trajectories are k of the S drawn ones (same noise), the
    weights are normalised, humans outside the cluster carry the cluster's weights."""
    B, H, S, k = 6, 5, 20, 8
    hist, rh = random_histories(B, H, seed=3, spread=1.5)
    noise = torch.randn(B, S, H, 8, 2, generator=torch.Generator().manual_seed(2)).cuda()
    f_all = forecaster("rand", B, H, S=S, k=S, step=5)
    f_all.set_history(torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    fc_all = f_all.predict(B, noise=noise)[0].cpu().numpy()
    _, n_in, ped, inc = (t.cpu().numpy() for t in f_all.encode(B))
    f_k = forecaster("rand", B, H, S=S, k=k, step=5)
    f_k.set_history(torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    fc_k, lw_k = (t.cpu().numpy() for t in f_k.predict(B, noise=noise))
    for b in range(B):
        ids = list(ped[b, :int(n_in[b])])
        picked = []
        for j in range(k):
            d = np.abs(fc_all[b][ids] - fc_k[b][ids][:, j:j + 1]).max(axis=(0, 2, 3))
            assert d.min() == 0.0, (b, j)
            picked.append(int(d.argmin()))
        assert len(set(picked)) == k
        assert abs(np.exp(lw_k[b][ids[0]]).sum() - 1.0) <= 1e-5
        for h in range(H):
            assert np.array_equal(lw_k[b][h], lw_k[b][ids[0]])
            if not inc[b, h]:
                assert np.array_equal(fc_k[b][h], fc_all[b][h][:k])       # constant-velocity rows do not depend on the sample


def test_kde_path_against_reference_golden_case():
    """Reference run with num_ret 8 < 20 drawn: its totals tie exactly (uniform weights log(1/8) in the golden file), so WHICH
    8 samples it returns is torch.argsort's tie order; we check the weights and that our 8 are drawn samples of the same noise."""
    tag = "rand_h4_kde"
    hist, rh = G[tag + "_hist"][..., :2], G[tag + "_robot_hist"][..., :2]
    H, n_draw, n_ret, step = (int(v) for v in G[tag + "_cfg"])
    ids_in = list(G[tag + "_ids_in"])
    nz = _noise_from_xT(G[tag + "_xT"], len(ids_in), n_draw, H)
    f = forecaster("rand", 1, H, S=n_draw, k=n_ret, step=step)
    fc, lw = f.predict_host(hist[None], rh[None], nz)
    ref_fc, ref_lw = G[tag + "_forecasts"], G[tag + "_logw"]
    assert np.max(np.abs(ref_lw - np.log(1.0 / n_ret))) <= 1e-6          # the golden weights are the degenerate uniform ones
    assert np.max(np.abs(lw[0] - ref_lw)) <= 1e-5
    assert np.array_equal(fc[0][list(G[tag + "_ids_out"])], ref_fc[list(G[tag + "_ids_out"])])
    f_all = forecaster("rand", 1, H, S=n_draw, k=n_draw, step=step)
    fc_all, _ = f_all.predict_host(hist[None], rh[None], nz)
    for j in range(n_ret):          # every reference trajectory is one of our S drawn ones (within the bf16 tolerance)
        d = np.abs(fc_all[0][ids_in] - ref_fc[ids_in][:, j:j + 1]).max(axis=(0, 2, 3))
        assert d.min() <= 2e-2, (j, d.min())


def test_mpc_ingest_is_bit_exact():
    B, H, S = 5, 7, 20
    hist, rh = random_histories(B, H, seed=21)
    f = forecaster("rand", B, H, S=S, step=2)
    f.set_history(torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    fc, lw = f.predict(B)                                 # library noise
    resh, wts, goals, vpref = (t.cpu().numpy() for t in f.ingest(fc, lw, horiz=4))
    fc, lw = fc.cpu().numpy(), lw.cpu().numpy()
    for b in range(B):
        r, w, g, v = PO.mpc_ingest(fc[b], lw[b], dt=0.25, horiz=4, joint=True)
        assert np.array_equal(resh[b], r) and np.array_equal(wts[b], w) and np.array_equal(goals[b], g) and np.array_equal(vpref[b], v)


def test_history_ring_push_semantics():
    B, H = 3, 4
    f = forecaster("rand", B, H, S=4, step=2)
    f.reset_history()
    rng = np.random.default_rng(0)
    frames = rng.normal(size=(9, B, H + 1, 2))
    for k in range(9):
        hp = torch.from_numpy(frames[k, :, 1:].copy()).cuda()
        rp = torch.from_numpy(frames[k, :, 0].copy()).cuda()
        f.push(hp[..., 0].contiguous(), hp[..., 1].contiguous(), rp[:, 0].contiguous(), rp[:, 1].contiguous())
    fc, _ = f.predict(B)
    # the rings hold frames 3..8: current pose = frame 8, and predicting from the explicit history gives the same result
    assert np.array_equal(fc.cpu().numpy()[:, :, 0, 0], frames[8, :, 1:])
    hist = np.ascontiguousarray(frames[3:9, :, 1:].transpose(1, 2, 0, 3)); rh = np.ascontiguousarray(frames[3:9, :, 0].transpose(1, 0, 2))
    f2_in = (torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    _, n1, p1, _ = f.encode(B)
    f.set_history(*f2_in)
    _, n2, p2, _ = f.encode(B)
    assert torch.equal(n1, n2) and torch.equal(p1, p2)


def test_library_noise_is_standard_normal_and_reproducible():
    from snb.jmid.forecaster import randn
    a = randn((1 << 20,), seed=7)
    b = randn((1 << 20,), seed=7)
    c = randn((1 << 20,), seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 5e-3 and abs(a.var().item() - 1.0) < 5e-3
    assert abs((a ** 4).mean().item() - 3.0) < 5e-2
    tail = randn((1 << 19,), seed=7, offset=(1 << 19) // 4)     # offset counts Philox counters (4 values each)
    assert torch.equal(tail, a[1 << 19:])


def test_drop_in_forecaster_sim_class():
    """The reference-facing class: configparser in, update_state_hists / predict_ret_best out (numpy fp64)."""
    from snb.jmid.forecaster import HumanTrajectoryForecasterSim
    tag = "rand_h5"
    H, n_draw, n_ret, step = (int(v) for v in G[tag + "_cfg"])
    cfg = configparser.RawConfigParser()
    cfg.read_dict({"env": {"time_step": "0.25"}, "sim": {"human_num": str(H)},
                   "human_trajectory_forecaster": {"past_num_frames": "6", "prediction_horizon": "8", "num_samples": str(n_ret),
                                                   "publish_freq": "0.08"}})
    _cache.clear()
    sim = HumanTrajectoryForecasterSim(cfg, {"num_samples": n_draw, "step_size": step, "joint_prediction": True},
                                       weights=(enc_weights("rand"), JO.make_random_weights(int(G["ddpm_seed"]))))

    class St:
        def __init__(self, p):
            self.position = (float(p[0]), float(p[1]))
    hist, rh = G[tag + "_hist"], G[tag + "_robot_hist"]
    for k in range(6):
        sim.update_state_hists(St(rh[k]), [St(hist[i, k]) for i in range(H)], hist[0, k, 2])
    ids_in = list(G[tag + "_ids_in"])
    fc, lw = sim.predict_ret_best(noise=_noise_from_xT(G[tag + "_xT"], len(ids_in), n_draw, H)[0])
    assert fc.dtype == np.float64 and fc.shape == (H, n_ret, 9, 2) and lw.shape == (H, n_ret)
    assert np.max(np.abs(fc - G[tag + "_forecasts"])) <= 2e-2
    assert np.array_equal(lw, G[tag + "_logw"])


def test_forecaster_sim_resamples_sparse_histories_like_the_reference():
    """Poses pushed every 0.5 s into a 0.25 s predictor: the reference's resample + linear interpolation (mid_sim_wrapper.py:283-298)
    fills the empty windows, so the forecasts must equal those of a second object fed the interpolated frames one time_step apart."""
    from snb.jmid.forecaster import HumanTrajectoryForecasterSim
    H, n_draw = 4, 6
    cfg = configparser.RawConfigParser()
    cfg.read_dict({"env": {"time_step": "0.25"}, "sim": {"human_num": str(H)},
                   "human_trajectory_forecaster": {"past_num_frames": "6", "prediction_horizon": "8", "num_samples": str(n_draw)}})
    w = (enc_weights("rand"), JO.make_random_weights(int(G["ddpm_seed"])))
    mk = lambda: HumanTrajectoryForecasterSim(cfg, {"num_samples": n_draw, "step_size": 5, "joint_prediction": True}, weights=w)
    sparse, dense = mk(), mk()

    class St:
        def __init__(self, p):
            self.position = (float(p[0]), float(p[1]))
    rng = np.random.default_rng(8)
    p0 = rng.uniform(-1.5, 1.5, (H + 1, 2)); v = rng.uniform(-0.6, 0.6, (H + 1, 2)); acc = rng.uniform(-0.3, 0.3, (H + 1, 2))
    pose = lambda t: p0 + v * t + 0.5 * acc * t * t                     # curved paths: interpolated frames differ from the true ones
    ts_sparse = [0.0, 0.5, 1.0, 1.5]                                      # -> windows 0, .25, ..., 1.5: 7 rows, the newest 6 are kept
    for t in ts_sparse:
        P = pose(t)
        sparse.update_state_hists(St(P[H]), [St(P[i]) for i in range(H)], t)
    for t in np.arange(0.25, 1.5 + 1e-9, 0.25):
        k = int(t // 0.5)
        P = pose(t) if abs(t - 0.5 * round(t / 0.5)) < 1e-12 else pose(0.5 * k) + (pose(0.5 * (k + 1)) - pose(0.5 * k)) * 0.5
        dense.update_state_hists(St(P[H]), [St(P[i]) for i in range(H)], float(t))
    nz = np.random.default_rng(9).standard_normal((n_draw, H, 8, 2)).astype(np.float32)
    f1, l1 = sparse.predict_ret_best(noise=nz)
    f2, l2 = dense.predict_ret_best(noise=nz)
    assert np.array_equal(f1, f2) and np.array_equal(l1, l2)
    # poses recorded faster than time_step leave fewer than 6 windows in the six retained raw entries (the reference's quirk): refused loudly
    fast = mk()
    for k in range(8):
        P = pose(0.05 * k)
        fast.update_state_hists(St(P[H]), [St(P[i]) for i in range(H)], 0.05 * k)
    from snb import _capi
    with pytest.raises(_capi.SnbError, match="after resampling"):
        fast.predict_ret_best(noise=nz)


IG = np.load(f"{GOLDEN}/ingest_cases.npz")


@pytest.mark.parametrize("tag", ["jmid_h10_k20", "jmid_h3_k15", "imid_h5_k8"])
def test_mpc_state_vector_matches_reference_function(tag):
    """snb_pred_mpc_pack vs the output of the reference's own convert_to_mpc_state_vector (sicnav_acados.py:222-289, executed by
    oracle/gen_golden.py ingest): bit-exact except sin / cos of the heading (CUDA libm vs glibc: <= 1 ulp)."""
    from snb import _capi
    joint = bool(IG[tag + "_joint"])
    humans, goals, w = IG[tag + "_humans"], IG[tag + "_goals"], IG[tag + "_weights"]
    H, k = humans.shape[0], w.shape[-1]
    B = 3
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape))).cuda()
    nx = 10 + (6 * H + k if joint else (6 + k) * H)
    state = torch.zeros(B, nx, dtype=torch.float64, device="cuda"); theta = torch.zeros(B, H, dtype=torch.float64, device="cuda")
    t_r, t_h, t_g, t_w = dev(IG[tag + "_robot"]), dev(humans), dev(goals), dev(w)        # kept alive across the call
    _capi.check(_capi.lib.snb_pred_mpc_pack(_capi.ptr(t_r), _capi.ptr(t_h), _capi.ptr(t_g), _capi.ptr(t_w),
                                            None, None, None, B, H, k, 8, 4, int(joint), 0, 0, _capi.ptr(state), _capi.ptr(theta), None,
                                            _capi.stream_ptr()), "mpc_pack")
    ref = IG[tag + "_val"]
    got = state.cpu().numpy()
    for b in range(B):
        assert np.array_equal(np.delete(got[b], [2, 3]), np.delete(ref, [2, 3]))
        assert np.max(np.abs(got[b][2:4] - ref[2:4])) <= 2.3e-16
    _, th = PO.mpc_state_vector(IG[tag + "_robot"], humans, goals, w, joint)
    assert np.max(np.abs(theta.cpu().numpy()[0] - th)) <= 4.5e-16 and theta[0, 1].item() == 0.0


def test_stage_parameter_packing_is_bit_exact():
    """Per-stage solver parameters (sicnav_acados.py:1389-1413) from the device-resident ingest outputs, vs the oracle restatement."""
    B, H, S, horiz = 4, 6, 20, 4
    hist, rh = random_histories(B, H, seed=33)
    f = forecaster("rand", B, H, S=S, step=2)
    f.set_history(torch.from_numpy(hist).cuda(), torch.from_numpy(rh).cuda())
    fc, lw = f.predict(B)
    resh, wts, goals, vpref = f.ingest(fc, lw, horiz=horiz)
    rng = np.random.default_rng(1)
    robot = rng.uniform(-2, 2, (B, 9)); humans = rng.uniform(-2, 2, (B, H, 4)); prefix = rng.normal(size=(B, horiz + 1, 11)); stat = rng.normal(size=(3, 4))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    state, theta, params = f.mpc_pack(t(robot), t(humans), goals, wts, resh=resh, horiz=horiz, stage_prefix=t(prefix), static_obs=t(stat))
    state, theta, params = state.cpu().numpy(), theta.cpu().numpy(), params.cpu().numpy()
    resh_h, w_h, g_h = resh.cpu().numpy(), wts.cpu().numpy(), goals.cpu().numpy()
    for b in range(B):
        assert np.array_equal(params[b], PO.stage_params(resh_h[b], horiz, prefix[b], stat))
        v, th = PO.mpc_state_vector(robot[b], humans[b], g_h[b], w_h[b], joint=True)
        assert np.array_equal(np.delete(state[b], [2, 3]), np.delete(v, [2, 3])) and np.max(np.abs(state[b][2:4] - v[2:4])) <= 2.3e-16
    # without prefix / static obstacles, and the horizon check (forecasts_reshaped keeps horiz+1 <= T frames)
    _, _, p2 = f.mpc_pack(t(robot), t(humans), goals, wts, resh=resh, horiz=horiz)
    assert np.array_equal(p2.cpu().numpy()[1], PO.stage_params(resh_h[1], horiz))


def test_history_bootstrap_from_the_environment_state_log():
    """reset_scenario_values (sicnav_acados.py:1163-1182): after reset with starts_moving = 10 the forecaster rings are filled from
    env.states[-7:-1]; the environment keeps that log on the device (snb_env_log_push, crowd_sim_plus.py:1175-1181)."""
    import bench
    from snb.env import CrowdSimPlusBatch
    B, H = 5, 4
    cfg = configparser.RawConfigParser()
    cfg.read_string(bench.ENV_CFG.format(H=H))
    env = CrowdSimPlusBatch(B, "cuda")
    env.configure(cfg)
    env.freeze_done = False
    log = []
    real_launch = env._launch

    def spy(action, active, *a, **k):           # records what the reference's self.states.append would hold: positions BEFORE the step
        s = env.state
        log.append(np.concatenate([np.stack([s.px.cpu().numpy(), s.py.cpu().numpy()], -1),
                                   np.stack([s.rpx.cpu().numpy(), s.rpy.cpu().numpy()], -1)[:, None]], 1))
        return real_launch(action, active, *a, **k)
    env._launch = spy
    env.reset('test', test_cases=np.arange(B))
    act = torch.zeros(B, 2, dtype=torch.float64, device="cuda"); act[:, 1] = 0.7
    for _ in range(3):
        env.step(act)
    assert env.n_logged == len(log) == 13
    f = forecaster("rand", B, H, S=4, step=2)
    f.reset_history()
    f.bootstrap_history(env.state_log, (env.n_logged - 1) % env.LOG_DEPTH)
    _, n1, p1, _ = f.encode(B)
    c1 = f.encode(B)[0].clone()
    for b in range(B):
        hist_b, rob_b = PO.bootstrap_history([l[b] for l in log])
        if b == 0:
            hist_all = np.zeros((B, H, 6, 2)); rob_all = np.zeros((B, 6, 2))
        hist_all[b], rob_all[b] = hist_b, rob_b
    f.set_history(torch.from_numpy(hist_all).cuda(), torch.from_numpy(rob_all).cuda())
    c2, n2, p2, _ = f.encode(B)
    assert torch.equal(n1, n2) and torch.equal(p1, p2) and torch.equal(c1, c2)
