"""CrowdSimPlusBatch -- B independent CrowdSimPlus environments stepped by ONE kernel launch.

The reference simulator (crowd_sim_plus/envs/crowd_sim_plus.py) is a single-environment gym.Env whose `step`
loops over Python Human objects.  This class keeps the reference's configuration surface (`configure(config)`
reads the same INI keys, `reset(phase, test_case)` builds the same seeded scenes, `step(action)` has the same
semantics incl. `starts_moving` warm-up, clamp, collision / frozen / goal / time-out flags and rewards) but holds
the state of all environments as fp64 SoA arrays in HBM (snb.state.CrowdStateSoA) and advances them with
snb_env_step.  There is no reference counterpart for the batching itself (SURVEY 8b "Python-side drop-ins").
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from . import _capi, scenario
from .policy.policy_factory import policy_factory
from .state import CrowdStateSoA, Obstacles


class CrowdSimPlusBatch:
    LOG_DEPTH = 7          # past_num_frames + 1 states: what the predictor's history bootstrap reads (sicnav_acados.py:1163-1182)

    def __init__(self, num_envs, device="cuda"):
        self.B = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.SnbError("CrowdSimPlusBatch needs a CUDA device (snb has no CPU path)")
        self.config = None
        self.state = None
        self.obstacles = None
        self.static_obstacles = []
        self.human_policy = None
        self.robot_kinematics = "holonomic"
        self.robot_visible = True
        self.freeze_done = True

    # ------------------------------------------------------------------ configuration
    def configure(self, config):
        """Same keys as CrowdSimPlus.configure (crowd_sim_plus.py:58-197), non-SB3 branch only: the SB3 branch whitelists rewards
        per RL model and adds angular / linear smoothness terms (crowd_sim_plus.py:69-75, 94-100) that the kernel's reward table
        does not have -- those configurations are refused instead of silently returning different rewards."""
        if config.getboolean('env', 'SB3', fallback=False):
            raise NotImplementedError("CrowdSimPlusBatch: [env] SB3 = True (per-model reward whitelist, RL observation spaces) is not implemented")
        if config.getboolean('env', 'occlusion', fallback=False):
            raise NotImplementedError("occlusion")           # as the reference (crowd_sim_plus.py:79-80)
        for key in ('angular_smoothness_factor', 'linear_smoothness_factor'):
            if config.has_option('reward', key) and float(config.get('reward', key)) != 0.0:
                raise NotImplementedError(f"CrowdSimPlusBatch: [reward] {key} (action-smoothness penalty) is not implemented")
        self.config = config
        self.time_limit = config.getfloat('env', 'time_limit')
        self.time_step = config.getfloat('env', 'time_step')
        self.randomize_attributes = config.getboolean('env', 'randomize_attributes')
        rewards = {k: float(v) for k, v in dict(config.items('reward')).items()}
        if "discomfort_dist" in rewards and "discomfort_penalty_factor" in rewards:
            rewards["discomfort"] = True
        else:
            rewards["discomfort_dist"] = 0.2
            rewards["discomfort"] = False
        rewards.setdefault("timeout", -1.0)
        rewards.setdefault("success_reward", 1.0)
        rewards.setdefault("collision_penalty", -1.0)
        rewards.setdefault("wall_collision_penalty", -1.0)
        rewards.setdefault("freezing_penalty", -1.0)
        self.rewards = rewards
        pol = config.get('humans', 'policy')
        if pol not in ('orca', 'orca_plus', 'sfm'):
            raise NotImplementedError(pol)
        self.case_capacity = {'train': np.iinfo(np.uint32).max - 2000, 'val': 1000, 'test': 1000}
        self.case_size = {'train': np.iinfo(np.uint32).max - 2000, 'val': config.getint('env', 'val_size'),
                          'test': config.getint('env', 'test_size')}
        self.train_val_sim = config.get('sim', 'train_val_sim')
        self.test_sim = config.get('sim', 'test_sim')
        self.square_width = config.getfloat('sim', 'square_width')
        self.circle_radius = config.getfloat('sim', 'circle_radius')
        self.rect_width = config.getfloat('sim', 'rect_width')
        self.rect_height = config.getfloat('sim', 'rect_height')
        self.starts_moving = config.getint('sim', 'starts_moving', fallback=0)
        self.human_num = config.getint('sim', 'human_num')
        # Human.__init__ / Agent.__init__ (agent_plus.py:12-28, human_plus.py:6-17)
        self.human_radius = config.getfloat('humans', 'radius')
        self.human_v_pref = config.getfloat('humans', 'v_pref')
        self.human_policy = policy_factory[pol]()
        try:
            self.human_policy.configure(config, 'humans')
        except Exception:   # reference swallows configure errors for humans (quirk q6)
            pass
        self.human_policy.time_step = self.time_step
        self.robot_radius = config.getfloat('robot', 'radius')
        self.robot_v_pref = config.getfloat('robot', 'v_pref')
        self.robot_visible = config.getboolean('robot', 'visible')

    def set_robot_kinematics(self, kinematics):
        assert kinematics in ("holonomic", "unicycle")
        self.robot_kinematics = kinematics

    # ------------------------------------------------------------------ cfg structs
    def _policy_cfg(self):
        from .policy._device_policy import policy_cfg
        return policy_cfg(self.human_policy, self.human_policy._KIND)

    def _door_cfg(self):
        d = self._door
        if d is None or self.sim_env not in scenario.DOOR_RULES or len(self.static_obstacles) == 0:
            return _capi.DoorCfg(enabled=0)
        return _capi.DoorCfg(enabled=1, **d)

    def _reward_cfg(self):
        r = self.rewards
        return _capi.RewardCfg(success_reward=r["success_reward"], timeout=r["timeout"], collision_penalty=r["collision_penalty"],
                               wall_collision_penalty=r["wall_collision_penalty"], freezing_penalty=r["freezing_penalty"],
                               discomfort=int(bool(r["discomfort"])), has_progress=int("progress_factor" in r),
                               discomfort_dist=r["discomfort_dist"], discomfort_penalty_factor=r.get("discomfort_penalty_factor", 0.0),
                               progress_factor=r.get("progress_factor", 0.0), time_limit=self.time_limit)

    # ------------------------------------------------------------------ reset
    def reset(self, phase='test', test_cases=None):
        """Builds env b from test case `test_cases[b]` (default b) with the reference's seeding
        (default_rng(offset + case), crowd_sim_plus.py:658-664), runs the `starts_moving` warm-up steps with a zero robot
        action (:709-720).  Returns the observation dict.
        The scenes are generated by snb_scene_reset, one thread per environment consuming the PCG64 stream of numpy's
        default_rng in the reference's draw order.  test_cases[0] == -1 selects the reference's fixed 3-human debug layout
        (:676-682) for every environment."""
        assert phase in ('train', 'val', 'test')
        self.phase = phase
        self.sim_env = self.test_sim if phase == 'test' else self.train_val_sim
        cases = np.arange(self.B) if test_cases is None else np.asarray(test_cases).reshape(self.B)
        p = scenario.SceneParams(self.circle_radius, self.rect_width, self.rect_height, self.human_radius, self.human_v_pref,
                                 self.robot_radius, self.rewards["discomfort_dist"], self.randomize_attributes)
        debug_case = bool(len(cases) and cases[0] == -1)
        H = 3 if debug_case else self.human_num
        kin = _capi.KIN_HOLONOMIC if self.robot_kinematics == "holonomic" else _capi.KIN_UNICYCLE
        st = CrowdStateSoA(self.B, H, 1, self.device, kin, self.robot_visible)
        if self.sim_env != 'circle_crossing' and self.sim_env not in scenario.HALLWAY_RULES:
            raise ValueError("Rule doesn't exist (square_crossing is broken in the reference, quirk q9)")
        segs, door = scenario.static_obstacles(self.sim_env, p)
        self._door = door
        self.static_obstacles = [[(s[0], s[1]), (s[2], s[3])] for s in segs]
        self.obstacles = Obstacles(segs) if len(segs) else None
        pol_name = self.config.get('humans', 'policy')
        if self.sim_env in scenario.HALLWAY_RULES and pol_name not in ('orca_plus', 'sfm'):
            raise RuntimeError("In hallway scenarios, human policy must be orca_plus or sfm, no other human policies supported due to "
                               "static obstacles")           # generate_hallway_human, crowd_sim_plus.py:524-525
        # the reference builds fresh Human / policy objects on every reset, so is_bottleneck starts False each time (:448-449)
        if hasattr(self.human_policy, 'is_bottleneck') or pol_name == 'sfm':
            self.human_policy.is_bottleneck = bool(self.sim_env == 'hallway_bottleneck' and pol_name == 'sfm')
        if not debug_case:
            cap = {'val': self.case_capacity['val'], 'test': self.case_capacity['test']}
            offset = {"train": cap["val"] + cap["test"], "val": 0, "test": cap["val"]}[phase]
            seeds = torch.from_numpy((cases.astype(np.int64) + offset).astype(np.uint64).view(np.int64)).to(self.device)
            scfg = _capi.SceneCfg(rule=_capi.SCENE_CIRCLE_CROSSING if self.sim_env == 'circle_crossing' else _capi.SCENE_HALLWAY,
                                  randomize_attributes=int(self.randomize_attributes), circle_radius=self.circle_radius,
                                  rect_width=self.rect_width, rect_height=self.rect_height, human_radius=self.human_radius,
                                  human_v_pref=self.human_v_pref, robot_radius=self.robot_radius, discomfort_dist=self.rewards["discomfort_dist"])
            segs_dev = torch.from_numpy(np.ascontiguousarray(segs, np.float64).reshape(-1)).to(self.device) if len(segs) else None
            self.reset_draws = torch.zeros(self.B, dtype=torch.int32, device=self.device)
            dc = self._door_cfg()
            cs = st.cstruct()
            _capi.check(_capi.lib.snb_scene_reset(C.byref(scfg), C.byref(dc), C.byref(cs), _capi.ptr(seeds), _capi.ptr(segs_dev), len(segs),
                                                  _capi.ptr(self.reset_draws), _capi.stream_ptr()), "snb_scene_reset")
            if int(self.reset_draws.min().item()) < 0:
                raise _capi.SnbError("scene reset: rejection sampling gave up (over-crowded scene: too many humans for this layout)")
        else:
            if self._door_cfg().enabled:
                raise NotImplementedError("the debug layout (test case -1) is only built for rules without a door goal")
            lay = scenario.debug_layout(p)          # px, py, gx, gy, v_pref, theta
            rep = lambda c: np.broadcast_to(lay[:, c], (self.B, H))
            st.load_numpy(px=rep(0), py=rep(1), gx=rep(2), gy=rep(3), fgx=rep(2), fgy=rep(3), vpref=rep(4), theta=rep(5),
                          radius=np.full((self.B, H), self.human_radius))
            st.ex_px.fill_(0.0); st.ex_py.fill_(-self.circle_radius); st.ex_radius.fill_(self.robot_radius)
            st.rgx.fill_(0.0); st.rgy.fill_(self.circle_radius); st.rtheta.fill_(np.pi / 2)
        self.state = st
        self.active = torch.ones(self.B, dtype=torch.uint8, device=self.device)
        self.reward = torch.zeros(self.B, dtype=torch.float64, device=self.device)
        self.dmin = torch.zeros(self.B, dtype=torch.float64, device=self.device)
        self.flags = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._zero_action = torch.zeros(self.B, 2, dtype=torch.float64, device=self.device)
        self._cfgs = (self._policy_cfg(), self._door_cfg(), self._reward_cfg())
        # `self.states` of the reference (crowd_sim_plus.py:1175-1181) as a ring of the last LOG_DEPTH pre-step positions
        self.state_log = torch.zeros(self.B, self.LOG_DEPTH, H + 1, 2, dtype=torch.float64, device=self.device) if self.LOG_DEPTH else None
        self.n_logged = 0
        if self.starts_moving > 0:
            st.global_time.fill_(-self.starts_moving * self.time_step)
            for _ in range(self.starts_moving):
                self._launch(self._zero_action, None)
        st.prev_dist.copy_(torch.hypot(st.rpx - st.rgx, st.rpy - st.rgy))
        self.check_status()         # a device capacity overflow during the warm-up steps must not go unnoticed
        return self.observation()

    # ------------------------------------------------------------------ step
    def _launch(self, action, active, stream=None, nbr=None, nbr_cnt=None):
        # the argument block is rebuilt only when the state object changes (reset): a step is then one ctypes call
        if getattr(self, "_argc_state", None) is not self.state:
            pc, dc, rc = self._cfgs
            self._argc = (C.byref(pc), C.byref(dc), C.byref(rc), self.state.cstruct())
            self._argc_state = self.state
            self._argc_tail = (_capi.ptr(self.reward), _capi.ptr(self.dmin), _capi.ptr(self.flags))
            self._argc_status = _capi.ptr(self.status)
        pc, dc, rc, st = self._argc
        if self.state_log is not None:        # states.append of this step rides in the same launch
            slot = self.n_logged % self.LOG_DEPTH
            self.n_logged += 1
            _capi.check(_capi.lib.snb_env_step_logged(pc, dc, rc, C.byref(st),
                                                      self.obstacles.handle if self.obstacles is not None else None,
                                                      _capi.ptr(action), _capi.ptr(active), *self._argc_tail,
                                                      _capi.ptr(nbr), _capi.ptr(nbr_cnt), self._argc_status,
                                                      _capi.ptr(self.state_log), self.LOG_DEPTH, slot,
                                                      _capi.stream_ptr(stream)), "snb_env_step_logged")
            return
        _capi.check(_capi.lib.snb_env_step(pc, dc, rc, C.byref(st),
                                           self.obstacles.handle if self.obstacles is not None else None,
                                           _capi.ptr(action), _capi.ptr(active), *self._argc_tail,
                                           _capi.ptr(nbr), _capi.ptr(nbr_cnt), self._argc_status,
                                           _capi.stream_ptr(stream)), "snb_env_step")

    def step(self, robot_action, stream=None, nbr=None, nbr_cnt=None):
        """robot_action: [B,2] fp64 CUDA tensor, (vx,vy) or (v,r).  Returns (reward[B], done[B] bool, flags[B]) on the
        device; state advances in place.  With `freeze_done`, environments that finished are frozen: the kernel skips them and
        their reward / flags read 0 from then on (so summing rewards over an episode does not re-count the terminal reward).
        A device-side capacity overflow (more than 32 ORCA lines / obstacle neighbours) is raised by check_status(), which
        reset() and rollouts call; a single step does not synchronise."""
        a = robot_action
        if not (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64 and a.is_contiguous()):
            a = torch.as_tensor(np.asarray(a, np.float64).reshape(self.B, 2)).to(self.device)
        with torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext():
            self._launch(a, self.active if self.freeze_done else None, stream, nbr, nbr_cnt)
            done = self.flags >= _capi.F_DONE            # F_DONE is the highest flag bit: one compare instead of and + ne
            if self.freeze_done:
                self.active &= (~done).to(torch.uint8)
        return self.reward, done, self.flags

    def what_if(self, robot_actions, stream=None):
        """`step(action, update=False)` (crowd_sim_plus.py:1025, :1239-1255) for A candidate robot actions per environment in
        one launch: robot_actions [B, A, 2] fp64 CUDA tensor.  Returns (reward [B,A], done [B,A] bool, flags [B,A],
        next_humans [B,H,4] = get_next_observable_state of every human, next_robot [B,A,2]); the state is not modified.
        This is the look-ahead SARL_input_complete / RGL_multistep_input_complete run once per discrete action (:797-866)."""
        a = robot_actions
        if not (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64 and a.is_contiguous()):
            a = torch.as_tensor(np.asarray(a, np.float64)).to(self.device).contiguous()
        assert a.dim() == 3 and a.shape[0] == self.B and a.shape[2] == 2
        A = int(a.shape[1])
        H = self.state.H
        reward = torch.empty(self.B, A, dtype=torch.float64, device=self.device)
        dmin = torch.empty(self.B, A, dtype=torch.float64, device=self.device)
        flags = torch.zeros(self.B, A, dtype=torch.int32, device=self.device)
        next_h = torch.empty(self.B, H, 4, dtype=torch.float64, device=self.device)
        next_r = torch.empty(self.B, A, 2, dtype=torch.float64, device=self.device)
        pc, dc, rc = self._cfgs
        st = self.state.cstruct()
        _capi.check(_capi.lib.snb_env_whatif(C.byref(pc), C.byref(dc), C.byref(rc), C.byref(st),
                                             self.obstacles.handle if self.obstacles is not None else None, _capi.ptr(a), A, None,
                                             _capi.ptr(reward), _capi.ptr(dmin), _capi.ptr(flags), _capi.ptr(next_h), _capi.ptr(next_r),
                                             _capi.ptr(self.status), _capi.stream_ptr(stream)), "snb_env_whatif")
        self.whatif_dmin = dmin
        return reward, (flags & _capi.F_DONE) != 0, flags, next_h, next_r

    def step_host(self, robot_action_np):
        """Host-buffer form used for end-to-end timing: H2D(action) -> step -> D2H(observation, reward, flags)."""
        a = torch.from_numpy(np.ascontiguousarray(robot_action_np, np.float64).reshape(self.B, 2)).to(self.device, non_blocking=True)
        reward, done, flags = self.step(a)
        ob = self.observation_host()
        return ob, reward.cpu().numpy(), done.cpu().numpy(), flags.cpu().numpy()

    def observation(self):
        """`ob = [human.get_observable_state() ...]` (crowd_sim_plus.py:1231) as device tensors [B,H]."""
        s = self.state
        return dict(px=s.px, py=s.py, vx=s.vx, vy=s.vy, radius=s.radius)

    def observation_host(self):
        s = self.state
        return torch.stack([s.px, s.py, s.vx, s.vy, s.radius], -1).cpu().numpy()

    def check_status(self):
        if int(self.status.item()) != 0:
            raise _capi.SnbError("device capacity exceeded during a crowd step (ORCA lines / obstacle neighbours)")
