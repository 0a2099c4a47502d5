"""CPU: the oracles against the reference's outputs with the SHIPPED JMID checkpoint at the C4 shape (tests/golden/ckpt_c4_cases.npz,
made by oracle/gen_golden.py ckpt), using the exported tensors tests/golden/ckpt_jmid_epoch121.npz -- this pins (a) the export,
(b) the fp32 oracle on trained weights at 1600 tokens x 20 iterations, which is what the GPU tests of tests/test_ckpt_gpu.py lean on."""
import os

import numpy as np
import pytest

import jmid_oracle as JO
import predictor_oracle as PO
import ref_shims
from golden_util import GOLDEN

torch = pytest.importorskip("torch")
C4 = np.load(f"{GOLDEN}/ckpt_c4_cases.npz")


def _weights():
    from snb.jmid.weights import load_checkpoint
    return load_checkpoint(f"{GOLDEN}/ckpt_jmid_epoch121.npz")


def test_product_synthetic_weights_equal_the_oracle_factories():
    from snb.jmid import weights as W
    a, b = W.synthetic_ddpm(5), JO.make_random_weights(5)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    a, b = W.synthetic_encoder(9), PO.make_random_encoder_weights(9)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.skipif(not ref_shims.have_reference(), reason="the .pt lives under /root/reference")
def test_export_equals_the_reference_checkpoint_file():
    from snb.jmid.weights import load_checkpoint
    enc, ddpm = _weights()
    enc2, ddpm2 = load_checkpoint(os.path.join(ref_shims.REF, "sicnav_diffusion/JMID/MID/checkpoints/sim_inference_checkpoints",
                                               "sim_gen_sicnav_p_midjp_cvg_epoch121.pt"))
    assert all(torch.equal(enc[k], enc2[k]) for k in enc) and all(torch.equal(ddpm[k], ddpm2[k]) for k in ddpm)


def test_oracle_noise_net_and_sampler_on_shipped_weights_c4_shape():
    _, w = _weights()
    tag = "ckpt_h10_dense"
    ctx, xT = torch.from_numpy(C4[tag + "_ctx"]), torch.from_numpy(C4[tag + "_xT"])
    A = ctx.shape[0]; S = xT.shape[0] // A
    betas = w["var_sched.betas"]
    with torch.no_grad():
        for t in (100, 55, 5):
            e = JO.net_forward(w, xT, betas[[t] * (A * S)], ctx.repeat(S, 1), joint=True)
            assert np.max(np.abs(e.numpy() - C4[f"{tag}_eps{t}"])) <= 1e-4, t
        out = JO.sample(w, ctx, xT, step=20, joint=True)
    assert np.max(np.abs(out.numpy() - C4[tag + "_vel"])) <= 1e-3


@pytest.mark.parametrize("tag", ["ckpt_h10", "ckpt_h3_shipped"])
def test_oracle_predict_ret_best_on_shipped_weights(tag):
    enc, w = _weights()
    hist, rh = C4[tag + "_hist"], C4[tag + "_robot_hist"]
    H, n_draw, n_ret, step = (int(v) for v in C4[tag + "_cfg"])
    with torch.no_grad():
        fc, lw, ctx = PO.predict_ret_best(enc, w, hist, rh, torch.from_numpy(C4[tag + "_xT"]), n_draw, n_ret, step)
    assert np.max(np.abs(ctx.numpy() - C4[tag + "_ctx"])) <= 2e-6
    ids_out = list(C4[tag + "_ids_out"])
    assert np.array_equal(fc[ids_out], C4[tag + "_forecasts"][ids_out])
    if n_ret == n_draw:
        assert np.max(np.abs(fc - C4[tag + "_forecasts"])) <= 1e-3
        assert np.array_equal(lw, C4[tag + "_logw"])
    else:
        assert np.max(np.abs(lw - C4[tag + "_logw"])) <= 1e-3
