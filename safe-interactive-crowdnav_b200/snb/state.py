"""SoA crowd state resident in HBM (the layout include/snb.h::SnbCrowdState points into).

All arrays are fp64 CUDA tensors: humans [B,H], extras/robot [B,E], per-env [B].  Replaces the per-agent Python
objects of crowd_sim_plus/envs/utils/{agent,human,robot}_plus.py for the batched simulator.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi

HUMAN_FIELDS = ("px", "py", "vx", "vy", "theta", "gx", "gy", "fgx", "fgy", "vpref", "radius", "human_time")
EXTRA_FIELDS = ("ex_px", "ex_py", "ex_vx", "ex_vy", "ex_radius")
ENV_FIELDS = ("rtheta", "rgx", "rgy", "global_time", "prev_dist")


class CrowdStateSoA:
    def __init__(self, B, H, E=1, device="cuda", robot_kinematics=_capi.KIN_HOLONOMIC, robot_visible=True):
        self.B, self.H, self.E = int(B), int(H), int(E)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.SnbError("CrowdStateSoA lives in HBM: a CUDA device is required (snb has no CPU path)")
        self.robot_kinematics = robot_kinematics
        self.robot_visible = robot_visible
        # one allocation per group keeps every field 16-byte aligned for the kernel's TMA bulk copies
        nh = ((self.B * self.H + 1) // 2) * 2
        ne = ((self.B * max(self.E, 1) + 1) // 2) * 2
        nb = ((self.B + 1) // 2) * 2
        self._hbuf = torch.zeros(len(HUMAN_FIELDS), nh, dtype=torch.float64, device=self.device)
        self._ebuf = torch.zeros(len(EXTRA_FIELDS), ne, dtype=torch.float64, device=self.device)
        self._bbuf = torch.zeros(len(ENV_FIELDS), nb, dtype=torch.float64, device=self.device)
        for k, n in enumerate(HUMAN_FIELDS):
            setattr(self, n, self._hbuf[k, :self.B * self.H].view(self.B, self.H))
        for k, n in enumerate(EXTRA_FIELDS):
            setattr(self, n, self._ebuf[k, :self.B * self.E].view(self.B, self.E))
        for k, n in enumerate(ENV_FIELDS):
            setattr(self, n, self._bbuf[k, :self.B])

    # robot = extra 0
    @property
    def rpx(self):
        return self.ex_px[:, 0]

    @property
    def rpy(self):
        return self.ex_py[:, 0]

    @property
    def rvx(self):
        return self.ex_vx[:, 0]

    @property
    def rvy(self):
        return self.ex_vy[:, 0]

    def cstruct(self):
        s = _capi.CrowdState(B=self.B, H=self.H, E=self.E, n_obs_extras=self.E if self.robot_visible else 0,
                             robot_kinematics=self.robot_kinematics)
        for n in HUMAN_FIELDS + EXTRA_FIELDS + ENV_FIELDS:
            setattr(s, n, getattr(self, n).data_ptr())
        return s

    def load_numpy(self, **arrays):
        """Host -> HBM upload of named fields (numpy arrays of the field's shape)."""
        for n, a in arrays.items():
            t = getattr(self, n)
            t.copy_(torch.as_tensor(np.asarray(a, np.float64).reshape(t.shape)), non_blocking=False)

    def to_numpy(self, *names):
        return {n: getattr(self, n).detach().cpu().numpy().copy() for n in (names or HUMAN_FIELDS + EXTRA_FIELDS + ENV_FIELDS)}

    def clone(self):
        o = CrowdStateSoA(self.B, self.H, self.E, self.device, self.robot_kinematics, self.robot_visible)
        o._hbuf.copy_(self._hbuf); o._ebuf.copy_(self._ebuf); o._bbuf.copy_(self._bbuf)
        return o


class Obstacles:
    """Static line segments shared by all environments; wraps SnbObstacles (host BSP build + device upload)."""

    def __init__(self, segments):
        segs = np.ascontiguousarray(np.asarray(segments, np.float64).reshape(-1, 4))
        self.segments = segs
        self._h = C.c_void_p()
        _capi.check(_capi.lib.snb_obstacles_create(C.byref(self._h), segs.ctypes.data_as(C.POINTER(C.c_double)), len(segs)),
                    "snb_obstacles_create")

    @property
    def handle(self):
        return self._h

    def __len__(self):
        return len(self.segments)

    def vertices(self):
        n = _capi.lib.snb_obstacles_num_vertices(self._h)
        out = np.zeros((n, 7), np.float32)
        buf = (C.c_float * 7)()
        for i in range(n):
            _capi.check(_capi.lib.snb_obstacles_get_vertex(self._h, i, buf), "snb_obstacles_get_vertex")
            out[i] = list(buf)
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        lib = getattr(_capi, "lib", None)      # None while the interpreter shuts down
        if h and lib is not None:
            lib.snb_obstacles_destroy(h)
            self._h = None
