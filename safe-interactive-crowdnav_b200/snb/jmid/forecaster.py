"""JMID predictor on the device: history rings -> clustering / scene graph / context encoder -> batched DDIM denoiser ->
integration, KDE top-k, forecasts and log-weights, MPC ingest (snb_pred_* in include/snb.h).

Two faces over the same C ABI:
  * `ForecasterBatch`               B environments at once, device tensors in and out (the data-parallel form);
  * `HumanTrajectoryForecasterSim`  drop-in for sicnav_diffusion/JMID/mid_sim_wrapper.py:207-509 (B = 1, the reference's
                                    constructor arguments, `update_state_hists`, `predict_ret_best` returning numpy fp64).
torch tensors only hold weights and I/O buffers; every computation is in libsnb.so (no CPU path).
"""
import ctypes as C
import threading

import numpy as np
import torch

from .. import _capi
from .denoiser import JmidDenoiser

from .weights import ATT, ENC_MODULES as _ENC, load_checkpoint, resolve_model_path  # noqa: F401  (load_checkpoint re-exported)

ENC_MODULES = {k: v[0] for k, v in _ENC.items()}


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


def read_mid_config(mid_config_file):
    """`MID.__init__` accepts a yaml PATH or an EasyDict (mid.py:83-89); the simulator passes the path
    "sicnav_diffusion/JMID/test_time_configs/mid_jp.yaml" (sicnav_acados.py:998).  Returns a plain dict."""
    if mid_config_file is None:
        return {}
    if isinstance(mid_config_file, (str, bytes)) or hasattr(mid_config_file, "__fspath__"):
        import os
        path = os.fspath(mid_config_file)
        if not os.path.isfile(path):
            alt = os.path.join(os.environ.get("SNB_REFERENCE", "/root/reference"), path)
            if os.path.isfile(alt):
                path = alt
        with open(path) as f:
            text = f.read()
        try:
            import yaml
            return dict(yaml.safe_load(text))
        except ImportError:
            return _parse_flat_yaml(text)
    if isinstance(mid_config_file, dict):
        return dict(mid_config_file)
    return dict(vars(mid_config_file))


def _parse_flat_yaml(text):
    """The test-time configs are flat `key: scalar | [list]` files; enough of YAML for them when PyYAML is absent."""
    out = {}
    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        if not line or ":" not in line:
            continue
        k, v = (x.strip() for x in line.split(":", 1))
        if v.startswith("[") and v.endswith("]"):
            out[k] = [_yaml_scalar(x.strip()) for x in v[1:-1].split(",") if x.strip()]
        else:
            out[k] = _yaml_scalar(v)
    return out


def _yaml_scalar(v):
    low = v.lower()
    if low in ("true", "yes", "on"):
        return True
    if low in ("false", "no", "off"):
        return False
    if low in ("none", "null", "~", ""):
        return None if low != "none" else "None"     # PyYAML reads a bare `None` as the string 'None'
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v.strip("'\"")


class ForecasterBatch:
    """B environments x H humans.  `encoder` / `ddpm` as returned by load_checkpoint (or synthetic dicts of the same keys)."""

    def __init__(self, encoder, ddpm, max_envs, H, num_samples=20, num_ret=None, step_size=20, horizon=8, joint=True, dt=0.25,
                 radius=3.0, device="cuda", seed=0, precision="bf16"):
        if _capi.lib.snb_pred_create is None:
            raise _capi.SnbError("libsnb.so was built without the predictor")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.SnbError("ForecasterBatch needs a CUDA device (snb has no CPU path)")
        self.B, self.H, self.S, self.T = int(max_envs), int(H), int(num_samples), int(horizon)
        self.k = self.S if num_ret is None else int(num_ret)
        self.step_size, self.joint, self.dt, self.radius, self.seed = int(step_size), bool(joint), float(dt), float(radius), int(seed)
        self.denoiser = JmidDenoiser(ddpm, max_envs=self.B, A=self.H, S=self.S, T=self.T, joint=self.joint, device=device,
                                     precision=precision)
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

    def set_position_std(self, pos_std):
        """0 (default): positions are standardised by the attention radius like the reference (preprocessing.py:477-478);
        a positive value pins the scale (benchmarks that widen `radius` only to force A = H)."""
        _capi.check(_capi.lib.snb_pred_set_position_std(self._h, float(pos_std)), "snb_pred_set_position_std")

    def bootstrap_history(self, state_log, newest, stream=None):
        """reset_scenario_values (sicnav_acados.py:1163-1182): rings <- env.states[-7:-1]; state_log [B,L,H+1,2] fp64 CUDA ring
        (CrowdSimPlusBatch.state_log), `newest` = slot of the last logged state."""
        B, L = state_log.shape[0], state_log.shape[1]
        assert state_log.is_cuda and state_log.dtype == torch.float64 and state_log.is_contiguous() and tuple(state_log.shape) == (B, L, self.H + 1, 2)
        _capi.check(_capi.lib.snb_pred_bootstrap_history(self._h, _capi.ptr(state_log), L, int(newest), B, _capi.stream_ptr(stream)),
                    "snb_pred_bootstrap_history")

    def mpc_pack(self, robot, humans, goals, weights, resh=None, horiz=4, stage_prefix=None, static_obs=None, stream=None):
        """convert_to_mpc_state_vector (sicnav_acados.py:222-289) + the per-stage parameter vectors (:1389-1413).
        robot [B,9] = px,py,theta,lvel,omega,v_dot,omega_dot,gx,gy; humans [B,H,4] = px,py,vx,vy; goals / weights / resh from ingest().
        -> mpc_state [B, 10+nX_hums], human_theta [B,H], stage_params [B, horiz+1, n_prefix + 4*H*k + n_static] (None without resh)."""
        B = robot.shape[0]
        for t, shp in ((robot, (B, 9)), (humans, (B, self.H, 4)), (goals, (B, self.H, 2))):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == shp, (t.shape, shp)
        nx = 10 + (6 * self.H + self.k if self.joint else (6 + self.k) * self.H)
        state = torch.empty(B, nx, dtype=torch.float64, device=self.device)
        theta = torch.empty(B, self.H, dtype=torch.float64, device=self.device)
        n_prefix = 0 if stage_prefix is None else int(stage_prefix.shape[-1])
        n_static = 0 if static_obs is None else int(static_obs.numel())
        params = None
        if resh is not None:
            assert tuple(resh.shape) == (B, horiz + 1, self.H * self.k, 2), resh.shape
            if stage_prefix is not None:
                assert tuple(stage_prefix.shape) == (B, horiz + 1, n_prefix) and stage_prefix.is_contiguous()
            params = torch.empty(B, horiz + 1, n_prefix + 4 * self.H * self.k + n_static, dtype=torch.float64, device=self.device)
        _capi.check(_capi.lib.snb_pred_mpc_pack(_capi.ptr(robot), _capi.ptr(humans), _capi.ptr(goals), _capi.ptr(weights), _capi.ptr(resh),
                                                _capi.ptr(stage_prefix), _capi.ptr(static_obs), B, self.H, self.k, self.T, int(horiz),
                                                int(self.joint), n_prefix, n_static, _capi.ptr(state), _capi.ptr(theta), _capi.ptr(params),
                                                _capi.stream_ptr(stream)), "snb_pred_mpc_pack")
        return state, theta, params

    def __del__(self):
        h = getattr(self, "_h", None)
        lib = getattr(_capi, "lib", None)      # None while the interpreter shuts down
        if h and lib is not None:
            lib.snb_pred_destroy(h)
            self._h = None


def kde_topk(pos, k):
    """get_most_likely_samples on the device: pos [B,S,A,T,2] fp32 CUDA -> (sel [B,k] int32, logw [B,k] fp64)."""
    B, S, A, T, _ = pos.shape
    assert pos.is_cuda and pos.dtype == torch.float32 and pos.is_contiguous()
    sel = torch.zeros(B, k, dtype=torch.int32, device=pos.device)   # zero-filled: a NaN total (singular covariance) leaves valid indices
    lw = torch.zeros(B, k, dtype=torch.float64, device=pos.device)
    _capi.check(_capi.lib.snb_pred_kde_topk(_capi.ptr(pos), B, S, A, T, int(k), _capi.ptr(sel), _capi.ptr(lw), _capi.stream_ptr()),
                "snb_pred_kde_topk")
    return sel, lw


def randn(shape, seed, offset=0, device="cuda"):
    """Standard-normal fp32 tensor from the library's Philox generator (snb_pred_noise)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    _capi.check(_capi.lib.snb_pred_noise(_capi.ptr(out), out.numel(), int(seed), int(offset), _capi.stream_ptr()), "snb_pred_noise")
    return out


class _AutoEncoderFacade:
    """`MID.model` (models/autoencoder.py): only the member inference callers reach, `.diffusion`."""

    def __init__(self, diffusion):
        self.diffusion = diffusion


class _MidModelFacade:
    """What callers reach through `forecaster.mid_model` / `.model` in the reference (mid.py:73-104): `num_samples`, `config`, and
    `model.diffusion.sample_sicnav_inference(...)` (models/diffusion.py:478) bound to the CUDA denoiser (built on first use)."""

    def __init__(self, cfg, drawn, ddpm=None, joint=True, max_agents=10, device="cuda"):
        self.config = type("MidConfig", (dict,), {"__getattr__": lambda s_, k: s_[k]})(cfg)
        self.num_samples = drawn
        self.sicnav_inference = True
        self._ddpm, self._joint, self._max_agents, self._device = ddpm, joint, max_agents, device
        self._ae = None

    @property
    def model(self):
        if self._ae is None:
            from .diffusion import DiffusionTraj
            self._ae = _AutoEncoderFacade(DiffusionTraj(self._ddpm, joint=self._joint, max_agents=self._max_agents, device=self._device))
        return self._ae

    def eval(self, env=None, *a, **k):
        raise _capi.SnbError("snb predictor: MID.eval(Environment) has no counterpart (scenes are built on the device); "
                             "call HumanTrajectoryForecasterSim.predict() or mid_model.model.diffusion.sample_sicnav_inference(...)")


class HumanTrajectoryForecasterSim:
    """Drop-in for mid_sim_wrapper.HumanTrajectoryForecasterSim (B = 1), constructed exactly as the reference does
    (sicnav_acados.py:996-1000):  HumanTrajectoryForecasterSim(env_config=<RawConfigParser>, mid_config_file="<...>/mid_jp.yaml").

    env_config: [human_trajectory_forecaster] publish_freq / past_num_frames / prediction_horizon / num_samples, [env] time_step,
    [sim] human_num (mid_sim_wrapper.py:171-195).  mid_config_file: yaml path, mapping or attribute object with `model_path`,
    `diffnet` (JointPredictionTransformerConcatLinear = JMID, TransformerConcatLinear = iMID), `num_samples` (drawn), `step_size`,
    `sampling` (test_time_configs/mid_jp.yaml).  `weights=(encoder, ddpm)` bypasses the checkpoint file (tests).
    The history join / resampling / interpolation of mid_sim_wrapper.py:244-298 is applied (snb/jmid/history.py; the identity in the
    simulator).  Restrictions, stated: past_num_frames must be 6 (the shipped value) and 6 frames must be left after resampling;
    sampling must be "ddim"."""

    def __init__(self, env_config=None, mid_config_file=None, weights=None, device="cuda", seed=0):
        if env_config is None:
            raise _capi.SnbError("HumanTrajectoryForecasterSim: env_config is required (the reference's default path is a ROS file)")
        g = env_config
        self.publish_freq = g.getfloat("human_trajectory_forecaster", "publish_freq", fallback=0.0)
        self.time_step = g.getfloat("env", "time_step")
        self.num_hist_frames = g.getint("human_trajectory_forecaster", "past_num_frames")
        self.predict_horizon = g.getint("human_trajectory_forecaster", "prediction_horizon")
        self.num_ret_samples = g.getint("human_trajectory_forecaster", "num_samples")
        self.num_hums = g.getint("sim", "human_num")
        if self.num_hist_frames != 6:
            raise _capi.SnbError("snb predictor: past_num_frames must be 6 (the shipped configuration)")
        cfg = read_mid_config(mid_config_file)
        if str(cfg.get("sampling", "ddim")) != "ddim":
            raise _capi.SnbError("snb predictor: only sampling = ddim (the shipped test-time setting) is implemented")
        diffnet = cfg.get("diffnet")
        if diffnet is None:
            joint = bool(cfg.get("joint_prediction", True))
        elif diffnet in ("JointPredictionTransformerConcatLinear", "TransformerConcatLinear"):
            joint = diffnet == "JointPredictionTransformerConcatLinear"
        else:
            raise _capi.SnbError(f"snb predictor: unknown diffnet {diffnet!r}")
        enc, ddpm = weights if weights is not None else load_checkpoint(resolve_model_path(cfg["model_path"]))
        drawn = int(cfg.get("num_samples", 20))
        self.batch = ForecasterBatch(enc, ddpm, max_envs=1, H=self.num_hums, num_samples=drawn,
                                     num_ret=min(self.num_ret_samples, drawn), step_size=int(cfg.get("step_size", 20)),
                                     horizon=self.predict_horizon, joint=joint, dt=self.time_step, device=device, seed=seed)
        self.mid_model = self.model = _MidModelFacade(cfg, drawn, ddpm, joint, self.num_hums, device)
        self.mid_env = None          # the Trajectron Environment object has no counterpart: scenes are built on the device
        self.prev_states = [[] for _ in range(self.num_hums)]
        self.prev_robot_states = []
        self.prev_states_lock = threading.Lock()      # mid_sim_wrapper.py:174: the histories may be appended from another thread (ROS callback)

    def update_state_hists(self, robot_state, human_states, time_stamp):
        with self.prev_states_lock:
            for i in range(self.num_hums):
                self.prev_states[i].append([*human_states[i].position, time_stamp])
                if len(self.prev_states[i]) > self.num_hist_frames:
                    self.prev_states[i].pop(0)
            self.prev_robot_states.append([*robot_state.position, time_stamp])
            if len(self.prev_robot_states) > self.num_hist_frames:   # the reference keeps it unbounded but only uses the joined tail
                self.prev_robot_states.pop(0)

    def _histories(self):
        """The joined, resampled frame table of mid_sim_wrapper.py:244-298 (identity when the frames are time_step apart)."""
        from .history import resample_histories
        with self.prev_states_lock:                  # snapshot under the lock, like _gen_agent_df (mid_sim_wrapper.py:251-258)
            prev = [list(p) for p in self.prev_states]
            prev_robot = list(self.prev_robot_states)
        hum, rob = resample_histories(prev, prev_robot, self.time_step, self.num_hist_frames)
        if hum.shape[1] < self.num_hist_frames:
            raise _capi.SnbError(f"snb predictor: {hum.shape[1]} frames are left after resampling the histories to time_step "
                                 f"(poses recorded faster than time_step, or stamps missing from an agent); the device pipeline "
                                 f"needs past_num_frames = {self.num_hist_frames}")
        return hum[None], rob[None]

    def predict_ret_best(self, noise=None):
        """-> (forecasts [H, k, T+1, 2] float64, log-weights [H, k] float64), mid_sim_wrapper.py:482-509."""
        if not self.prev_states[0] or not self.prev_robot_states:
            raise _capi.SnbError("predict_ret_best: no history frames (call update_state_hists first)")
        hist, rob = self._histories()          # raises unless past_num_frames frames are left after the reference's resampling
        fc, lw = self.batch.predict_host(hist, rob, None if noise is None else noise[None])
        return fc[0], lw[0]

    def predict(self, noise=None):
        """mid_sim_wrapper.py:456-479: None until past_num_frames frames are held, then ALL drawn samples with the current pose
        prepended, [H, S, T+1, 2].  (The reference body calls convert_to_mid_state_env with one argument and unpacks two of its five
        results, so it raises as shipped; this is its evident intent.)"""
        if len(self.prev_states[0]) < self.num_hist_frames:
            return None
        k0 = self.batch.k
        self.batch.k = self.batch.S
        try:
            hist, rob = self._histories()
            fc, _ = self.batch.predict_host(hist, rob, None if noise is None else noise[None])
        finally:
            self.batch.k = k0
        return fc[0]

    def get_most_likely_samples(self, forecasts):
        """forecasts [S, A, T, 2] (torch or numpy) -> (top-k forecasts [A, k, T, 2], log-weights [A, k]) like
        mid_sim_wrapper.get_most_likely_samples(forecasts, model, num_ret_samples) (:14-169, :440-441)."""
        f = torch.as_tensor(forecasts, dtype=torch.float32, device=self.batch.device).contiguous()
        S, A = int(f.shape[0]), int(f.shape[1])
        k = min(self.num_ret_samples, S)
        sel, lw = kde_topk(f[None], k)
        top = f[sel[0].long()].permute(1, 0, 2, 3).contiguous()
        return top, lw[0].to(torch.float32)[None].expand(A, k)
