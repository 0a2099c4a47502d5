"""CPU tests: the denoiser oracle (oracle/jmid_oracle.py) against outputs of the reference's own
models/diffusion.py (tests/golden/jmid_cases.npz, made by oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

import jmid_oracle as JO
import ref_shims
from golden_util import GOLDEN

G = np.load(f"{GOLDEN}/jmid_cases.npz")


def test_variance_schedule_known_answers():
    betas, alphas, abar = JO.variance_schedule()
    assert np.allclose(abar.numpy(), G["jmid_alpha_bars"], atol=1e-7)
    assert np.allclose(betas.numpy(), G["jmid_betas"], atol=0)
    # SURVEY.md Appendix C.1
    kat = [0.078234, 0.100572, 0.127588, 0.159739, 0.197376, 0.240701, 0.289717, 0.344190, 0.403613, 0.467187, 0.533811,
           0.602105, 0.670436, 0.736985, 0.799815, 0.856969, 0.906565, 0.946903, 0.976559, 0.994471, 1.0]
    assert np.allclose(abar.numpy()[100::-5], kat, atol=2e-6)


@pytest.mark.parametrize("tag,joint", [("jmid", True), ("imid", False)])
@pytest.mark.parametrize("size", ["small", "c4"])
def test_oracle_matches_reference_module_with_seeded_weights(tag, joint, size):
    """Reference module loaded with make_random_weights(5) vs the oracle on the same weights (fp32 CPU both)."""
    w = JO.make_random_weights(int(G["rand_seed"]))
    ctx = torch.from_numpy(G[f"{tag}_rand_{size}_ctx"])
    xT = torch.from_numpy(G[f"{tag}_rand_{size}_xT"])
    A = ctx.shape[0]
    S = xT.shape[0] // A
    betas, _, _ = JO.variance_schedule()
    with torch.no_grad():
        e = JO.net_forward(w, xT, betas[[55] * (A * S)], ctx.repeat(S, 1), joint=joint)
        assert np.max(np.abs(e.numpy() - G[f"{tag}_rand_{size}_eps55"])) < 2e-4
        out = JO.sample(w, ctx, xT, step=int(G[f"{tag}_rand_{size}_steps"]), joint=joint)
    assert np.max(np.abs(out.numpy() - G[f"{tag}_rand_{size}_sample"])) < 2e-3


@pytest.mark.skipif(not ref_shims.have_reference(), reason="shipped checkpoints live under /root/reference")
@pytest.mark.parametrize("tag,joint,ckpt,diffnet", [
    ("jmid", True, "sim_gen_sicnav_p_midjp_cvg_epoch121.pt", "JointPredictionTransformerConcatLinear"),
    ("imid", False, "sim_gen_sicnav_p_mid_cvg_epoch169.pt", "TransformerConcatLinear")])
def test_oracle_matches_reference_with_shipped_checkpoint(tag, joint, ckpt, diffnet):
    _, ck = ref_shims.load_jmid_reference(ckpt, diffnet)
    w = {k[len("vel_predictor."):]: v.float() for k, v in ck["ddpm"].items() if k.startswith("vel_predictor.")}
    A, S, T = 2, 3, 8
    ctx = torch.linspace(-1, 1, A * 256).view(A, 256)
    x = torch.linspace(-1, 1, A * S * T * 2).view(A * S, T, 2)
    betas, _, _ = JO.variance_schedule()
    with torch.no_grad():
        e = JO.net_forward(w, x, betas[[100] * (A * S)], ctx.repeat(S, 1), joint=joint)
        assert np.max(np.abs(e.numpy() - G[f"{tag}_ckpt_kat_eps"])) < 1e-4
        smp = JO.sample(w, ctx, torch.zeros(A * S, T, 2), step=20, joint=joint)   # bestof=False -> x_T = 0
        assert np.max(np.abs(smp.numpy() - G[f"{tag}_ckpt_kat_sample"])) < 1e-3
        ctx2, xT = torch.from_numpy(G[f"{tag}_ckpt_ctx"]), torch.from_numpy(G[f"{tag}_ckpt_xT"])
        out = JO.sample(w, ctx2, xT, step=20, joint=joint)
        assert np.max(np.abs(out.numpy() - G[f"{tag}_ckpt_sample20"])) < 1e-3
    # SURVEY.md Appendix C.2 spot values
    if tag == "jmid":
        assert np.allclose(e[0].flatten()[:4].numpy(), [-0.833384, -0.442049, -0.771793, -0.385241], atol=2e-5)
        assert abs(float(e.abs().sum()) - 43.265472) < 2e-3
        assert abs(float(smp.abs().sum()) - 71.630913) < 5e-3
