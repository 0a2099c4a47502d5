"""State records of the reference plugin surface (crowd_sim_plus/envs/utils/state_plus.py:1-66).

Same constructor signatures and attribute names (px, py, vx, vy, radius, gx, gy, v_pref, theta, omega, position,
velocity, goal_position; JointState.self_state / human_states / static_obs), implemented as slotted dataclasses.
They exist for the B=1 `policy.predict(state)` call; the batched simulator never builds them (it keeps the SoA
arrays of snb.state.CrowdStateSoA in HBM).
"""
from dataclasses import dataclass, field
from typing import Any, List, Optional


@dataclass
class ObservableState:
    px: float
    py: float
    vx: float
    vy: float
    radius: float

    @property
    def position(self):
        return (self.px, self.py)

    @property
    def velocity(self):
        return (self.vx, self.vy)

    def as_tuple(self):
        return (self.px, self.py, self.vx, self.vy, self.radius)

    def __add__(self, other):          # reference: `other + (fields...)`
        return other + self.as_tuple()

    def __str__(self):
        return " ".join(str(x) for x in self.as_tuple())


@dataclass
class FullState:
    px: float
    py: float
    vx: float
    vy: float
    radius: float
    gx: float
    gy: float
    v_pref: float
    theta: float
    omega: Optional[float] = None

    @property
    def position(self):
        return (self.px, self.py)

    @property
    def velocity(self):
        return (self.vx, self.vy)

    @property
    def goal_position(self):
        return (self.gx, self.gy)

    def as_tuple(self):
        t = (self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)
        return t if self.omega is None else t + (self.omega,)

    def __add__(self, other):
        return other + self.as_tuple()

    def __str__(self):
        return " ".join(str(x) for x in (self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref,
                                         self.theta, self.omega))


@dataclass
class JointState:
    self_state: FullState
    human_states: List[Any]
    static_obs: List[Any] = field(default_factory=list)

    def __post_init__(self):
        assert isinstance(self.self_state, FullState)
        for s in self.human_states:
            assert isinstance(s, ObservableState)


@dataclass
class FullyObservableJointState:
    self_state: FullState
    human_states: List[Any]
    static_obs: List[Any] = field(default_factory=list)

    def __post_init__(self):
        assert isinstance(self.self_state, FullState)
        for s in self.human_states:
            assert isinstance(s, FullState)
