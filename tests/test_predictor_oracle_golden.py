"""CPU: oracle/predictor_oracle.py against tests/golden/predictor_cases.npz, which holds outputs of the REFERENCE's own
HumanTrajectoryForecasterSim.predict_ret_best (sicnav_diffusion/JMID/mid_sim_wrapper.py:482-509) run on CPU with the shipped
checkpoint (ckpt_* cases: encoder pinned; the 40 MB denoiser weights do not travel) and with seeded synthetic weights
(rand_* cases: the whole call pinned, incl. the KDE top-k branch).  Generator: oracle/gen_golden.py predictor."""
import numpy as np
import pytest

import jmid_oracle as JO
import predictor_oracle as PO
from golden_util import GOLDEN

torch = pytest.importorskip("torch")
G = np.load(f"{GOLDEN}/predictor_cases.npz")
CASES = ["ckpt_h5", "ckpt_h5_sparse", "rand_h5", "rand_h5_sparse", "rand_h10", "rand_h4_kde", "ckpt_h4_kde"]


def enc_weights(tag):
    if tag.startswith("ckpt"):
        return {k[9:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("ckpt_enc:")}
    return PO.make_random_encoder_weights(int(G["enc_seed"]))


@pytest.mark.parametrize("tag", CASES)
def test_cluster_split_and_context_match_reference(tag):
    hist, rh = G[tag + "_hist"], G[tag + "_robot_hist"]
    inp = PO.encoder_inputs(hist, rh)
    assert inp["ped_ids"] == list(G[tag + "_ids_in"])
    assert [i for i in range(hist.shape[0]) if not inp["in_cluster"][i]] == list(G[tag + "_ids_out"])
    ctx = PO.encode(enc_weights(tag), inp).numpy()
    assert ctx.shape == G[tag + "_ctx"].shape
    assert np.max(np.abs(ctx - G[tag + "_ctx"])) <= 2e-6


@pytest.mark.parametrize("tag", [c for c in CASES if c.startswith("rand")])
def test_predict_ret_best_matches_reference(tag):
    hist, rh = G[tag + "_hist"], G[tag + "_robot_hist"]
    H, n_draw, n_ret, step = (int(v) for v in G[tag + "_cfg"])
    with torch.no_grad():
        fc, lw, _ = PO.predict_ret_best(enc_weights(tag), JO.make_random_weights(int(G["ddpm_seed"])), hist, rh,
                                        torch.from_numpy(G[tag + "_xT"]), n_draw, n_ret, step)
    assert fc.shape == G[tag + "_forecasts"].shape == (H, n_ret, 9, 2)
    assert np.max(np.abs(fc - G[tag + "_forecasts"])) <= 1e-4
    assert np.max(np.abs(lw - G[tag + "_logw"])) <= 1e-4
    # current pose first, constant-velocity rows for the humans outside the cluster
    assert np.array_equal(fc[:, :, 0], np.repeat(hist[:, -1, None, :2], n_ret, axis=1))
    for h in G[tag + "_ids_out"]:
        v = (hist[h, -1, :2] - hist[h, -2, :2]) / 0.25
        assert np.allclose(fc[h, :, 1:], hist[h, -1, :2] + v * 0.25 * np.arange(1, 9)[:, None], atol=1e-12)


def test_mpc_ingest_layout():
    """sicnav_acados.py:1645-1667 restated; the acados policy cannot be imported here (parity of this packing: unpinned,
    checked against an independent einops/numpy evaluation of the same expressions)."""
    einops = pytest.importorskip("einops")
    fc, lw = G["rand_h10_forecasts"], G["rand_h10_logw"]
    resh, w, goals, vpref = PO.mpc_ingest(fc, lw, dt=0.25, horiz=4, joint=True)
    f = fc[:, :, 1:, :]
    assert np.array_equal(resh, einops.rearrange(f, "h s t d -> t (h s) d")[:5])
    assert np.array_equal(w, lw[0, :])
    for h in range(fc.shape[0]):
        assert goals[h, 0] == np.mean(f[h, :, 0, 0]) and goals[h, 1] == np.mean(f[h, :, 0, 1])
        assert vpref[h] == np.max(np.linalg.norm(np.diff(f[h], axis=1), axis=2) / 0.25)


IG = np.load(f"{GOLDEN}/ingest_cases.npz")


@pytest.mark.parametrize("tag", ["jmid_h10_k20", "jmid_h3_k15", "imid_h5_k8"])
def test_mpc_state_vector_matches_reference_function(tag):
    """ingest_cases.npz holds outputs of the reference's own convert_to_mpc_state_vector (source cut out of sicnav_acados.py:222-289
    and executed, oracle/gen_golden.py ingest)."""
    val, theta = PO.mpc_state_vector(IG[tag + "_robot"], IG[tag + "_humans"], IG[tag + "_goals"], IG[tag + "_weights"],
                                     joint=bool(IG[tag + "_joint"]))
    assert np.array_equal(val, IG[tag + "_val"])
    assert theta[1] == 0.0 and np.all(np.isfinite(theta))


def test_stage_params_and_bootstrap_layout():
    rng = np.random.default_rng(3)
    H, k, horiz = 4, 3, 4
    resh = rng.normal(size=(horiz + 1, H * k, 2)); prefix = rng.normal(size=(horiz + 1, 7)); stat = rng.normal(size=(2, 4))
    p = PO.stage_params(resh, horiz, prefix, stat)
    assert p.shape == (horiz + 1, 7 + 4 * H * k + 8)
    for idx in range(horiz + 1):
        t = min(idx, horiz - 1)
        ref = np.hstack([prefix[idx], resh[t, :, 0], resh[t, :, 1], resh[t + 1, :, 0], resh[t + 1, :, 1], stat.reshape(-1)])
        assert np.array_equal(p[idx], ref)
    states = [rng.normal(size=(H + 1, 2)) for _ in range(10)]
    hist, rob = PO.bootstrap_history(states)
    assert np.array_equal(hist[:, 0], states[-7][:H]) and np.array_equal(hist[:, -1], states[-2][:H]) and np.array_equal(rob[-1], states[-2][H])
    with pytest.raises(IndexError):
        PO.bootstrap_history(states[:6])
