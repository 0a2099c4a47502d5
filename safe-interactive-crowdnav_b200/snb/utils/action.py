"""Action records of the reference plugin surface (crowd_sim_plus/envs/utils/action.py:1-4): same names and
field order, so `ActionXY(*v)` / `action.vx` written against the reference keep working."""
import collections

ActionXY = collections.namedtuple("ActionXY", "vx vy")    # holonomic: world-frame velocity
ActionRot = collections.namedtuple("ActionRot", "v r")    # unicycle: speed, heading change over the step
