from .denoiser import JmidDenoiser, weights_struct  # noqa: F401
from .diffusion import DiffusionTraj  # noqa: F401
