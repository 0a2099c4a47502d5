"""JMID predictor on the device: history rings -> clustering / scene graph / context encoder -> batched DDIM denoiser ->
integration, KDE top-k, forecasts and log-weights, MPC ingest (snb_pred_* in include/snb.h).

Two faces over the same C ABI:
  * `ForecasterBatch`               B environments at once, device tensors in and out (the data-parallel form);
  * `HumanTrajectoryForecasterSim`  drop-in for sicnav_diffusion/JMID/mid_sim_wrapper.py:207-509 (B = 1, the reference's
                                    constructor arguments, `update_state_hists`, `predict_ret_best` returning numpy fp64).
torch tensors only hold weights and I/O buffers; every computation is in libsnb.so (no CPU path).
"""
import ctypes as C
import pickle

import numpy as np
import torch

from .. import _capi
from .denoiser import JmidDenoiser

ENC_MODULES = {
    "node_history": "PEDESTRIAN/node_history_encoder",
    "edge_ped": "PEDESTRIAN->PEDESTRIAN/edge_encoder",
    "edge_robot": "PEDESTRIAN->JRDB_ROBOT/edge_encoder",
}
ATT = "PEDESTRIAN/edge_influence_encoder"


def encoder_struct(enc, device):
    """{"<module>/<param>": tensor} (the flattened `checkpoint["encoder"]`) -> (SnbEncoderWeights, keepalive)."""
    keep = []

    def dev(name):
        t = enc[name].detach().to(device=device, dtype=torch.float32).contiguous()
        keep.append(t)
        return t.data_ptr()

    w = _capi.EncoderWeights()
    for field, mod in ENC_MODULES.items():
        l = getattr(w, field)
        l.w_ih = dev(f"{mod}/weight_ih_l0"); l.w_hh = dev(f"{mod}/weight_hh_l0")
        l.b_ih = dev(f"{mod}/bias_ih_l0"); l.b_hh = dev(f"{mod}/bias_hh_l0")
    w.att_w1 = dev(f"{ATT}/w1.weight"); w.att_w2 = dev(f"{ATT}/w2.weight"); w.att_v = dev(f"{ATT}/v.weight")
    return w, keep


class _TolerantUnpickler(pickle.Unpickler):
    """The reference pickles whole nn.Modules (`registrar.model_dict`, mid.py:1502-1505); their classes live in the reference
    tree.  Classes that cannot be imported are replaced by bare nn.Module subclasses: only the parameters are needed."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return type(name, (torch.nn.Module,), {"__module__": module})


class _TolerantPickle:
    __name__ = "pickle"
    Unpickler = _TolerantUnpickler
    load = staticmethod(lambda f, **kw: _TolerantUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def load_checkpoint(path):
    """Reads a reference checkpoint {"encoder": ModuleDict, "ddpm": state_dict} (mid.py:1231-1232, 1291) ->
    (encoder dict "<module>/<param>" -> tensor, ddpm state_dict)."""
    try:
        ck = torch.load(path, map_location="cpu", weights_only=False)
    except (ImportError, AttributeError, ModuleNotFoundError):
        ck = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
    enc = {}
    md = ck["encoder"]
    items = md.items() if hasattr(md, "items") else md._modules.items()
    for mod_name, mod in items:
        params = mod.named_parameters() if hasattr(mod, "named_parameters") else mod.items()
        for pname, t in params:
            enc[f"{mod_name}/{pname}"] = t.detach() if hasattr(t, "detach") else torch.as_tensor(t)
    return enc, ck["ddpm"]


class ForecasterBatch:
    """B environments x H humans.  `encoder` / `ddpm` as returned by load_checkpoint (or synthetic dicts of the same keys)."""

    def __init__(self, encoder, ddpm, max_envs, H, num_samples=20, num_ret=None, step_size=20, horizon=8, joint=True, dt=0.25,
                 radius=3.0, device="cuda", seed=0):
        if _capi.lib.snb_pred_create is None:
            raise _capi.SnbError("libsnb.so was built without the predictor")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.SnbError("ForecasterBatch needs a CUDA device (snb has no CPU path)")
        self.B, self.H, self.S, self.T = int(max_envs), int(H), int(num_samples), int(horizon)
        self.k = self.S if num_ret is None else int(num_ret)
        self.step_size, self.joint, self.dt, self.radius, self.seed = int(step_size), bool(joint), float(dt), float(radius), int(seed)
        self.denoiser = JmidDenoiser(ddpm, max_envs=self.B, A=self.H, S=self.S, T=self.T, joint=self.joint, device=device)
        w, keep = encoder_struct(encoder, self.device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.snb_pred_create(C.byref(self._h), C.byref(w), self.denoiser._h, self.B, self.H, _capi.stream_ptr()),
                        "snb_pred_create")
            torch.cuda.current_stream().synchronize()
        del keep

    # ---- history (update_state_hists) ----
    def push(self, human_px, human_py, robot_px, robot_py, stream=None):
        """fp64 CUDA tensors: human_p{x,y} [B,H], robot_p{x,y} [B]; appends one frame to the rings."""
        B = robot_px.shape[0]
        for t, shp in ((human_px, (B, self.H)), (human_py, (B, self.H)), (robot_px, (B,)), (robot_py, (B,))):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == shp, (t.shape, shp)
        _capi.check(_capi.lib.snb_pred_push_history(self._h, _capi.ptr(human_px), _capi.ptr(human_py), _capi.ptr(robot_px),
                                                    _capi.ptr(robot_py), B, _capi.stream_ptr(stream)), "snb_pred_push_history")

    def reset_history(self):
        _capi.check(_capi.lib.snb_pred_reset_history(self._h), "snb_pred_reset_history")

    def set_history(self, hist, robot_hist, stream=None):
        """hist [B,H,6,2], robot_hist [B,6,2] fp64 CUDA tensors, oldest frame first."""
        B = hist.shape[0]
        assert hist.is_cuda and hist.dtype == torch.float64 and hist.is_contiguous() and tuple(hist.shape) == (B, self.H, 6, 2)
        assert robot_hist.is_cuda and robot_hist.dtype == torch.float64 and robot_hist.is_contiguous() and tuple(robot_hist.shape) == (B, 6, 2)
        _capi.check(_capi.lib.snb_pred_set_history(self._h, _capi.ptr(hist), _capi.ptr(robot_hist), B, _capi.stream_ptr(stream)),
                    "snb_pred_set_history")

    # ---- encoder only ----
    def encode(self, B, stream=None):
        """-> ctx [B,H,256] fp32, n_in [B] int32, ped_ids [B,H] int32, in_cluster [B,H] uint8 (slot layout, see snb.h)."""
        ctx = torch.empty(B, self.H, 256, dtype=torch.float32, device=self.device)
        n_in = torch.empty(B, dtype=torch.int32, device=self.device)
        ped = torch.empty(B, self.H, dtype=torch.int32, device=self.device)
        inc = torch.empty(B, self.H, dtype=torch.uint8, device=self.device)
        _capi.check(_capi.lib.snb_pred_encode(self._h, B, self.radius, self.dt, _capi.ptr(ctx), _capi.ptr(n_in), _capi.ptr(ped),
                                              _capi.ptr(inc), _capi.stream_ptr(stream)), "snb_pred_encode")
        return ctx, n_in, ped, inc

    # ---- predict_ret_best ----
    def predict(self, B, noise=None, out=None, stream=None):
        """-> forecasts [B,H,k,T+1,2] fp64, logw [B,H,k] fp64 (CUDA).  noise: optional [B,S,H,T,2] fp32 CUDA tensor."""
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and \
                tuple(noise.shape) == (B, self.S, self.H, self.T, 2), noise.shape
        if out is None:
            fc = torch.zeros(B, self.H, self.k, self.T + 1, 2, dtype=torch.float64, device=self.device)
            lw = torch.zeros(B, self.H, self.k, dtype=torch.float64, device=self.device)
        else:
            fc, lw = out
        _capi.check(_capi.lib.snb_pred_predict(self._h, B, _capi.ptr(noise), self.seed, self.step_size, self.k, self.radius, self.dt,
                                               _capi.ptr(fc), _capi.ptr(lw), _capi.stream_ptr(stream)), "snb_pred_predict")
        return fc, lw

    def predict_host(self, hist_np, robot_hist_np, noise_np=None):
        """Host buffers in, host buffers out (the plugin call): hist [B,H,6,2], robot_hist [B,6,2] fp64."""
        hist_np = np.ascontiguousarray(hist_np, np.float64); robot_hist_np = np.ascontiguousarray(robot_hist_np, np.float64)
        B = hist_np.shape[0]
        assert hist_np.shape == (B, self.H, 6, 2) and robot_hist_np.shape == (B, 6, 2)
        fc = np.zeros((B, self.H, self.k, self.T + 1, 2), np.float64); lw = np.zeros((B, self.H, self.k), np.float64)
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        nz = None
        if noise_np is not None:
            noise_np = np.ascontiguousarray(noise_np, np.float32)
            assert noise_np.shape == (B, self.S, self.H, self.T, 2)
            nz = noise_np.ctypes.data_as(fp)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib.snb_pred_predict_host(self._h, hist_np.ctypes.data_as(dp), robot_hist_np.ctypes.data_as(dp), B, nz,
                                                        self.seed, self.step_size, self.k, self.radius, self.dt,
                                                        fc.ctypes.data_as(dp), lw.ctypes.data_as(dp)), "snb_pred_predict_host")
        return fc, lw

    # ---- MPC ingest (sicnav_acados.py:1645-1667) ----
    def ingest(self, forecasts, logw, horiz, stream=None):
        """-> forecasts_reshaped [B,min(T,horiz+1),H*k,2], weights [B,k] (joint) or [B,H,k], goals [B,H,2], v_pref [B,H] (fp64 CUDA)."""
        B = forecasts.shape[0]
        Tp = min(self.T, horiz + 1)
        resh = torch.empty(B, Tp, self.H * self.k, 2, dtype=torch.float64, device=self.device)
        wts = torch.empty((B, self.k) if self.joint else (B, self.H, self.k), dtype=torch.float64, device=self.device)
        goals = torch.empty(B, self.H, 2, dtype=torch.float64, device=self.device)
        vpref = torch.empty(B, self.H, dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib.snb_pred_ingest(_capi.ptr(forecasts), _capi.ptr(logw), B, self.H, self.k, self.T, int(horiz), self.dt,
                                              int(self.joint), _capi.ptr(resh), _capi.ptr(wts), _capi.ptr(goals), _capi.ptr(vpref),
                                              _capi.stream_ptr(stream)), "snb_pred_ingest")
        return resh, wts, goals, vpref

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _capi.lib.snb_pred_destroy(h)
            self._h = None


def kde_topk(pos, k):
    """get_most_likely_samples on the device: pos [B,S,A,T,2] fp32 CUDA -> (sel [B,k] int32, logw [B,k] fp64)."""
    B, S, A, T, _ = pos.shape
    assert pos.is_cuda and pos.dtype == torch.float32 and pos.is_contiguous()
    sel = torch.empty(B, k, dtype=torch.int32, device=pos.device)
    lw = torch.empty(B, k, dtype=torch.float64, device=pos.device)
    _capi.check(_capi.lib.snb_pred_kde_topk(_capi.ptr(pos), B, S, A, T, int(k), _capi.ptr(sel), _capi.ptr(lw), _capi.stream_ptr()),
                "snb_pred_kde_topk")
    return sel, lw


def randn(shape, seed, offset=0, device="cuda"):
    """Standard-normal fp32 tensor from the library's Philox generator (snb_pred_noise)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    _capi.check(_capi.lib.snb_pred_noise(_capi.ptr(out), out.numel(), int(seed), int(offset), _capi.stream_ptr()), "snb_pred_noise")
    return out


class HumanTrajectoryForecasterSim:
    """Drop-in for mid_sim_wrapper.HumanTrajectoryForecasterSim (B = 1).

    env_config: configparser with [human_trajectory_forecaster] past_num_frames / prediction_horizon / num_samples,
    [env] time_step, [sim] human_num (mid_sim_wrapper.py:171-195).  mid_config: mapping / attribute object with model_path,
    num_samples (drawn), step_size, joint_prediction (test_time_configs/mid_jp.yaml), or `weights=(encoder, ddpm)`."""

    def __init__(self, env_config, mid_config_file=None, weights=None, device="cuda", seed=0):
        g = env_config
        self.time_step = g.getfloat("env", "time_step")
        self.num_hist_frames = g.getint("human_trajectory_forecaster", "past_num_frames")
        self.predict_horizon = g.getint("human_trajectory_forecaster", "prediction_horizon")
        self.num_ret_samples = g.getint("human_trajectory_forecaster", "num_samples")
        self.num_hums = g.getint("sim", "human_num")
        if self.num_hist_frames != 6:
            raise _capi.SnbError("snb predictor: past_num_frames must be 6 (the shipped configuration)")
        cfg = {} if mid_config_file is None else (dict(mid_config_file) if isinstance(mid_config_file, dict) else dict(vars(mid_config_file)))
        enc, ddpm = weights if weights is not None else load_checkpoint(cfg["model_path"])
        drawn = int(cfg.get("num_samples", 20))
        self.batch = ForecasterBatch(enc, ddpm, max_envs=1, H=self.num_hums, num_samples=drawn,
                                     num_ret=min(self.num_ret_samples, drawn), step_size=int(cfg.get("step_size", 20)),
                                     horizon=self.predict_horizon, joint=bool(cfg.get("joint_prediction", True)), dt=self.time_step,
                                     device=device, seed=seed)
        self.prev_states = [[] for _ in range(self.num_hums)]
        self.prev_robot_states = []

    def update_state_hists(self, robot_state, human_states, time_stamp):
        for i in range(self.num_hums):
            self.prev_states[i].append([*human_states[i].position, time_stamp])
            if len(self.prev_states[i]) > self.num_hist_frames:
                self.prev_states[i].pop(0)
        self.prev_robot_states.append([*robot_state.position, time_stamp])
        if len(self.prev_robot_states) > self.num_hist_frames:   # the reference keeps it unbounded but only uses the joined tail
            self.prev_robot_states.pop(0)

    def predict_ret_best(self, noise=None):
        if len(self.prev_states[0]) < self.num_hist_frames:
            raise _capi.SnbError("predict_ret_best: fewer than past_num_frames history frames")
        hist = np.asarray(self.prev_states, np.float64)[None, :, :, :2]
        rob = np.asarray(self.prev_robot_states[-self.num_hist_frames:], np.float64)[None, :, :2]
        fc, lw = self.batch.predict_host(hist, rob, None if noise is None else noise[None])
        return fc[0], lw[0]
