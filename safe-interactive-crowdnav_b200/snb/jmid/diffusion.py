"""Drop-in for the inference entry of the reference's DiffusionTraj (sicnav_diffusion/JMID/MID/models/diffusion.py:478-541):

    traj, number_of_steps = diffusion.sample_sicnav_inference(num_points, context, sample, bestof, point_dim=2, flexibility=0.0,
                                                              ret_traj=False, sampling="ddim", step=20)

with the reference's argument names, order and return value, so the call site in AutoEncoder.generate_sicnav_inference
(models/autoencoder.py:17-47) binds to the CUDA denoiser unchanged.  The whole DDIM loop (all `step` iterations of the noise
network + update) is one C-ABI call, snb_jmid_denoise_agents; nothing leaves the device between iterations (the reference moves
every intermediate to the host, diffusion.py:532).
"""
import torch

from .. import _capi
from .denoiser import JmidDenoiser


class DiffusionTraj:
    """`net` / `var_sched` of the reference object are replaced by the `ddpm` state_dict of the checkpoint (the keys
    `net.*`, `var_sched.*` that DiffusionTraj.load_state_dict consumes).  joint = True: JointPredictionTransformerConcatLinear (JMID),
    False: TransformerConcatLinear (iMID).  Handles are created per (sample, num_points) on first use and kept."""

    def __init__(self, ddpm_state_dict, joint=True, max_agents=10, device="cuda", precision="bf16"):
        self._sd = ddpm_state_dict
        self.joint = bool(joint)
        self.max_agents = int(max_agents)
        self.device = torch.device(device)
        self.precision = precision
        self._handles = {}

    def _handle(self, S, T, A):
        if A > self.max_agents:
            self.max_agents = A
            self._handles.clear()
        key = (S, T)
        if key not in self._handles:
            self._handles[key] = JmidDenoiser(self._sd, max_envs=1, A=self.max_agents, S=S, T=T, joint=self.joint, device=self.device,
                                              precision=self.precision)
        return self._handles[key]

    @torch.no_grad()
    def sample_sicnav_inference(self, num_points, context, sample, bestof, point_dim=2, flexibility=0.0, ret_traj=False,
                                sampling="ddpm", step=100, with_constraints=True, dynamics=None, x_T=None):
        """context: [A, 256] (one row per agent of the cluster).  Returns (velocities [sample, A, num_points, 2], number_of_steps)
        exactly as diffusion.py:539-541.  `x_T` (extension, [sample * A, num_points, 2], row = s * A + a) injects the start noise
        instead of drawing it; bestof = False starts from zeros like the reference (diffusion.py:503-506).
        Not implemented, refused loudly: sampling = "ddpm" (the shipped test-time configs use ddim; diffusion.py:521-522), ret_traj,
        point_dim != 2, step values whose stride int(100 / step) does not divide 100 (the reference's loop then never reaches t = 0
        and its `traj[0]` raises KeyError).  step = 40 runs 50 iterations of stride 2, as the reference does."""
        if sampling != "ddim":
            raise _capi.SnbError(f"snb DiffusionTraj: sampling={sampling!r} is not implemented (ddim only)")
        if ret_traj:
            raise _capi.SnbError("snb DiffusionTraj: ret_traj=True is not implemented (intermediates stay on the device)")
        if point_dim != 2:
            raise _capi.SnbError("snb DiffusionTraj: point_dim must be 2")
        step = int(step)
        if step < 1 or step > 100:
            raise _capi.SnbError(f"snb DiffusionTraj: step={step} out of range")
        stride = int(100 / step)                                          # diffusion.py:508
        if 100 % stride != 0:
            raise _capi.SnbError(f"snb DiffusionTraj: step={step} gives stride {stride}, which does not divide the 100 diffusion steps")
        context = torch.as_tensor(context)
        if context.dim() != 2 or context.shape[1] != 256:
            raise _capi.SnbError(f"snb DiffusionTraj: context must be [A, 256], got {tuple(context.shape)}")
        A, S, T = int(context.shape[0]), int(sample), int(num_points)
        den = self._handle(S, T, A)
        ctx = context.to(self.device, torch.float32).contiguous()[None]
        if x_T is not None:
            x = torch.as_tensor(x_T).to(self.device, torch.float32).reshape(1, S * A, T, 2).contiguous()
        elif bestof:
            x = torch.randn(1, S * A, T, 2, device=self.device, dtype=torch.float32)
        else:
            x = torch.zeros(1, S * A, T, 2, device=self.device, dtype=torch.float32)
        out = den.denoise(ctx, x, n_steps=step)[0]                       # [S, A, T, 2]; the library applies the same int(100 / step)
        number_of_steps = S * (100 // stride + 1)
        return out.to(context.device) if context.device != out.device else out, number_of_steps
