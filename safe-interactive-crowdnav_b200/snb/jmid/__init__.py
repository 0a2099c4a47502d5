from .denoiser import JmidDenoiser, weights_struct  # noqa: F401
