"""GPU parity with the SHIPPED JMID checkpoint (sim_gen_sicnav_p_midjp_cvg_epoch121.pt), run with -m gpu, through the C ABI.

Weights: tests/golden/ckpt_jmid_epoch121.npz, the tensor export of the reference checkpoint (oracle/gen_golden.py ckpt).
Checker: tests/golden/ckpt_c4_cases.npz + predictor_cases.npz -- outputs of the REFERENCE's own HumanTrajectoryForecasterSim /
DiffusionTraj (fp32, CPU) with that checkpoint and injected x_T, at the C4 per-env shape (10 humans x 20 samples x 8 steps = 1600
tokens, 20 DDIM iterations), at a 5-of-10 cluster and at the shipped simulation setting (100 drawn, 2 iterations, 15 kept).

Arithmetic of the CUDA path: bf16 operands, fp32 accumulation / softmax / LayerNorm statistics / DDIM state.  Tolerances
(bf16 path <= 2e-2 m on positions; the reference computes in fp32):
  eps (one forward, |eps| up to 4.5)     <= 2.5e-2 absolute
  sampled velocities (|v| up to 1.7 m/s) <= 2e-2 m/s
  forecast positions                     <= 2e-2 m
The measured maxima are written to gpurun_out/ckpt_parity.json (quoted in DESIGN.md section 2).
"""
import configparser
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

C4 = np.load(f"{GOLDEN}/ckpt_c4_cases.npz")
PC = np.load(f"{GOLDEN}/predictor_cases.npz")
_w = {}
_measured = {}


def weights():
    if not _w:
        from snb.jmid.weights import load_checkpoint
        _w["enc"], _w["ddpm"] = load_checkpoint(f"{GOLDEN}/ckpt_jmid_epoch121.npz")
    return _w["enc"], _w["ddpm"]


def _record(key, val):
    _measured[key] = float(val)
    out = os.path.join(os.path.dirname(os.path.dirname(GOLDEN)), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "ckpt_parity.json"), "w") as f:
        json.dump(_measured, f, indent=1, sort_keys=True)


def _case(tag):
    return C4 if tag + "_hist" in C4.files else PC


def test_exported_checkpoint_is_what_the_encoder_golden_pinned():
    """predictor_cases.npz carries the shipped encoder tensors (recorded from the live reference module in round 1)."""
    enc, ddpm = weights()
    for k in PC.files:
        if k.startswith("ckpt_enc:"):
            assert np.array_equal(enc[k[9:]].numpy(), PC[k]), k
    assert ddpm["var_sched.alpha_bars"].shape == (101,) and ddpm["net.pos_emb.pe"].reshape(-1, 512).shape[0] >= 8


@pytest.mark.parametrize("tag,ts", [("ckpt_h10_dense", (100, 55, 5)), ("ckpt_h10", (55,)), ("ckpt_h3_shipped", (100,))])
def test_noise_net_forward_with_shipped_weights(tag, ts):
    """One forward of the noise network (diffusion.py:173-209) on (x_T, ctx) of the reference run."""
    from snb.jmid import JmidDenoiser
    ctx = torch.from_numpy(C4[tag + "_ctx"]); xT = torch.from_numpy(C4[tag + "_xT"])
    A = ctx.shape[0]; S = xT.shape[0] // A
    den = JmidDenoiser(weights()[1], max_envs=1, A=A, S=S, T=8, joint=True)
    for t in ts:
        e = den.eps(ctx[None].cuda(), xT[None].cuda(), t).cpu().numpy()[0]
        err = np.max(np.abs(e - C4[f"{tag}_eps{t}"]))
        _record(f"eps_t{t}_{tag}", err)
        assert err <= 2.5e-2, (t, err)


@pytest.mark.parametrize("tag", ["ckpt_h10_dense", "ckpt_h10", "ckpt_h3_shipped"])
def test_ddim_sampler_with_shipped_weights(tag):
    """sample_sicnav_inference (diffusion.py:478-541) with the reference's injected x_T: C4 shape 20 iterations; shipped 2 iterations."""
    from snb.jmid import JmidDenoiser
    ctx = torch.from_numpy(C4[tag + "_ctx"]); xT = torch.from_numpy(C4[tag + "_xT"])
    A = ctx.shape[0]; S = xT.shape[0] // A
    step = int(C4[tag + "_cfg"][3])
    den = JmidDenoiser(weights()[1], max_envs=2, A=A, S=S, T=8, joint=True)
    # two copies of the environment in one batch: both must equal the reference (and each other, bit for bit)
    out = den.denoise(ctx[None].repeat(2, 1, 1).cuda(), xT[None].repeat(2, 1, 1, 1).cuda(), n_steps=step).cpu().numpy()
    assert np.array_equal(out[0], out[1])
    err = np.max(np.abs(out[0] - C4[tag + "_vel"]))
    _record(f"vel_{tag}", err)
    assert err <= 2e-2, err


def _noise_from_xT(xT, A, S, H, T=8):
    nz = np.zeros((1, S, H, T, 2), np.float32)
    nz[0, :, :A] = xT.reshape(S, A, T, 2)
    return nz


@pytest.mark.parametrize("tag", ["ckpt_h10_dense", "ckpt_h10", "ckpt_h5", "ckpt_h5_sparse"])
def test_predict_ret_best_with_shipped_checkpoint(tag):
    """The whole call (histories -> clustering -> encoder -> 20-step denoise -> integrate -> forecasts) vs the reference's output."""
    from snb.jmid.forecaster import ForecasterBatch
    G = _case(tag)
    hist, rh = G[tag + "_hist"][..., :2], G[tag + "_robot_hist"][..., :2]
    H, n_draw, n_ret, step = (int(v) for v in G[tag + "_cfg"])
    enc, ddpm = weights()
    f = ForecasterBatch(enc, ddpm, max_envs=1, H=H, num_samples=n_draw, num_ret=n_ret, step_size=step)
    ids_in, ids_out = list(G[tag + "_ids_in"]), list(G[tag + "_ids_out"])
    f.set_history(torch.from_numpy(np.ascontiguousarray(hist[None])).cuda(), torch.from_numpy(np.ascontiguousarray(rh[None])).cuda())
    ctx, n_in, ped, _ = (t.cpu().numpy() for t in f.encode(1))
    assert int(n_in[0]) == len(ids_in) and list(ped[0, :len(ids_in)]) == ids_in
    assert np.max(np.abs(ctx[0, :len(ids_in)] - G[tag + "_ctx"])) <= 2e-5
    fc, lw = f.predict_host(hist[None], rh[None], _noise_from_xT(G[tag + "_xT"], len(ids_in), n_draw, H))
    ref_fc, ref_lw = G[tag + "_forecasts"], G[tag + "_logw"]
    assert np.array_equal(lw[0], ref_lw)
    assert np.array_equal(fc[0][:, :, 0], ref_fc[:, :, 0])
    assert np.array_equal(fc[0][ids_out], ref_fc[ids_out])
    err = np.max(np.abs(fc[0][ids_in] - ref_fc[ids_in]))
    _record(f"forecast_m_{tag}", err)
    assert err <= 2e-2, err


@pytest.mark.parametrize("tag", ["ckpt_h3_shipped", "ckpt_h4_kde"])
def test_kde_topk_path_with_shipped_checkpoint(tag):
    """num_ret < drawn (the shipped simulation setting keeps 15 of 100): the reference's per-step KDE totals tie exactly unless
    samples nearly coincide, so which k it returns is torch.argsort's tie order; checked: every returned reference trajectory is one
    of our drawn samples within tolerance, the log-weights match, the constant-velocity rows are bit-equal."""
    from snb.jmid.forecaster import ForecasterBatch
    G = _case(tag)
    hist, rh = G[tag + "_hist"][..., :2], G[tag + "_robot_hist"][..., :2]
    H, n_draw, n_ret, step = (int(v) for v in G[tag + "_cfg"])
    enc, ddpm = weights()
    ids_in, ids_out = list(G[tag + "_ids_in"]), list(G[tag + "_ids_out"])
    nz = _noise_from_xT(G[tag + "_xT"], len(ids_in), n_draw, H)
    f_all = ForecasterBatch(enc, ddpm, max_envs=1, H=H, num_samples=n_draw, num_ret=n_draw, step_size=step)
    fc_all, _ = f_all.predict_host(hist[None], rh[None], nz)
    del f_all
    f = ForecasterBatch(enc, ddpm, max_envs=1, H=H, num_samples=n_draw, num_ret=n_ret, step_size=step)
    fc, lw = f.predict_host(hist[None], rh[None], nz)
    ref_fc, ref_lw = G[tag + "_forecasts"], G[tag + "_logw"]
    assert fc.shape[1:] == ref_fc.shape
    assert np.array_equal(fc[0][ids_out], ref_fc[ids_out])
    # The KDE bandwidth is 0.01 .. 0.1 m (mid_sim_wrapper.py:60-62): position differences of a few mm between the bf16 denoiser and
    # the fp32 reference move the per-sample log-likelihoods by O(0.1).  Tolerance on the (ascending) log-weights: 0.15; they stay
    # normalised and sorted.  (precision="fp32x" tightens this, see test_fp32x_gpu.py.)
    lw_err = np.max(np.abs(lw[0] - ref_lw))
    _record(f"kde_logw_{tag}", lw_err)
    assert lw_err <= 0.15, lw_err
    assert np.all(np.diff(lw[0][ids_in[0]]) >= -1e-6) and abs(np.exp(lw[0][ids_in[0]]).sum() - 1.0) <= 1e-5
    worst = 0.0
    for j in range(n_ret):
        d = np.abs(fc_all[0][ids_in] - ref_fc[ids_in][:, j:j + 1]).max(axis=(0, 2, 3))
        worst = max(worst, float(d.min()))
    _record(f"forecast_m_{tag}", worst)
    assert worst <= 2e-2, worst


def test_reference_literal_constructor_call(tmp_path, monkeypatch):
    """sicnav_acados.py:996-1000: HumanTrajectoryForecasterSim(env_config=self.env.config, mid_config_file="<...>/mid_jp.yaml"),
    then update_state_hists x 6 and predict_ret_best (:1173-1182, :1641-1644).  The yaml written here carries the keys the reference's
    test_time_configs/mid_jp.yaml sets for inference; its model_path names the shipped .pt, which is resolved to the exported tensors."""
    from snb.jmid.forecaster import HumanTrajectoryForecasterSim
    tag = "ckpt_h10_dense"
    H, n_draw, n_ret, step = (int(v) for v in C4[tag + "_cfg"])
    ydir = tmp_path / "sicnav_diffusion" / "JMID" / "test_time_configs"
    ydir.mkdir(parents=True)
    (ydir / "mid_jp.yaml").write_text(
        "seed: 0\ntime: False  # comment\neval_mode: True\nmethod: mid_jp\nmaximum_history_length: 5\nprediction_horizon: 8\n"
        "model_path: sicnav_diffusion/JMID/MID/checkpoints/sim_inference_checkpoints/sim_gen_sicnav_p_midjp_cvg_epoch121.pt\n"
        "joint_prediction: True\ndiffnet: JointPredictionTransformerConcatLinear\nencoder_dim: 256\ntf_layer: 3\n"
        f"num_samples: {n_draw}\nstep_size: {step}  # step size of each iteration of reverse diffusion\nsampling: ddim\n"
        "override_attention_radius: []\n")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("SNB_REFERENCE", str(tmp_path / "no_reference_here"))
    cfg = configparser.RawConfigParser()
    cfg.read_dict({"env": {"time_step": "0.25"}, "sim": {"human_num": str(H)},
                   "human_trajectory_forecaster": {"past_num_frames": "6", "prediction_horizon": "8", "num_samples": str(n_ret),
                                                   "publish_freq": "0.08"}})
    sim = HumanTrajectoryForecasterSim(env_config=cfg, mid_config_file="sicnav_diffusion/JMID/test_time_configs/mid_jp.yaml")
    assert sim.num_hist_frames == 6 and sim.predict_horizon == 8 and sim.num_ret_samples == n_ret and sim.num_hums == H
    assert sim.mid_model.num_samples == n_draw and sim.model is sim.mid_model and sim.mid_model.config.step_size == step
    assert sim.predict() is None                                  # not enough history yet (mid_sim_wrapper.py:457-458)

    class St:
        def __init__(self, p):
            self.position = (float(p[0]), float(p[1]))
    hist, rh = C4[tag + "_hist"], C4[tag + "_robot_hist"]
    for k in range(6):
        sim.update_state_hists(St(rh[k]), [St(hist[i, k]) for i in range(H)], hist[0, k, 2])
    nz = _noise_from_xT(C4[tag + "_xT"], H, n_draw, H)[0]
    fc, lw = sim.predict_ret_best(noise=nz)
    assert fc.dtype == np.float64 and fc.shape == (H, n_ret, 9, 2) and lw.dtype == np.float64 and lw.shape == (H, n_ret)
    assert np.max(np.abs(fc - C4[tag + "_forecasts"])) <= 2e-2 and np.array_equal(lw, C4[tag + "_logw"])
    allf = sim.predict(noise=nz)
    assert allf.shape == (H, n_draw, 9, 2) and np.array_equal(allf, fc)      # k = S here: predict() returns every drawn sample
    sim.num_ret_samples = 8                                       # the method form of the KDE top-k (mid_sim_wrapper.py:440-441)
    pos2 = torch.from_numpy(np.ascontiguousarray(np.cumsum(C4[tag + "_vel"][:, :2], axis=2) * 0.25))
    top, w = sim.get_most_likely_samples(pos2)
    assert tuple(top.shape) == (2, 8, 8, 2) and tuple(w.shape) == (2, 8) and abs(float(w[0].exp().sum()) - 1.0) < 1e-4
    flat = pos2.reshape(n_draw, -1)
    assert all(bool((flat == top[:, j].cpu().reshape(1, -1)).all(1).any()) for j in range(8))     # the kept ones are drawn samples
    # the inner entry the reference's AutoEncoder calls (models/autoencoder.py:33-44 -> diffusion.py:478), through the same object chain:
    # forecaster.mid_model (MID) .model (AutoEncoder) .diffusion (DiffusionTraj); shipped weights, the reference's x_T and velocities
    ctx = torch.from_numpy(C4[tag + "_ctx"])
    traj, nsteps = sim.mid_model.model.diffusion.sample_sicnav_inference(8, ctx.cuda(), n_draw, True, point_dim=2, flexibility=0.0, ret_traj=False,
                                                                         sampling="ddim", step=step, x_T=torch.from_numpy(C4[tag + "_xT"]))
    assert tuple(traj.shape) == (n_draw, H, 8, 2) and nsteps == n_draw * (step + 1)
    assert np.max(np.abs(traj.cpu().numpy() - C4[tag + "_vel"])) <= 2e-2


def test_step_size_must_divide_the_diffusion_steps():
    """The reference raises KeyError on traj[0] when int(100/step) does not divide 100 (diffusion.py:507-537); here SNB_EINVAL."""
    from snb import _capi
    from snb.jmid import JmidDenoiser
    den = JmidDenoiser(weights()[1], max_envs=1, A=2, S=2, T=8, joint=True)
    ctx = torch.zeros(1, 2, 256, device="cuda"); xT = torch.zeros(1, 4, 8, 2, device="cuda")
    for bad in (3, 30, 7, 8):
        with pytest.raises(_capi.SnbError, match="does not divide"):
            den.denoise(ctx, xT, n_steps=bad)
    for ok in (1, 2, 4, 5, 10, 20, 25, 50, 100, 34, 51):          # int(100/34) = 2, int(100/51) = 1
        den.denoise(ctx, xT, n_steps=ok)
