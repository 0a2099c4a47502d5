"""snb -- B200-native drop-in for the data-parallel hot path of sepsamavi/safe-interactive-crowdnav.

Host code is Python (like the reference); all arithmetic of the hot path runs in hand-written sm_100a CUDA kernels
reached through the C ABI of libsnb.so (include/snb.h) via ctypes.  torch is used only to own device memory,
streams and (multi-GPU) the process group.  There is no CPU fallback: importing `snb._capi` fails loudly when
libsnb.so has not been built, and every call fails when no CUDA device is present.
"""
__version__ = "0.1.0"
