"""String -> policy class registry with the keys of crowd_sim_plus/envs/policy/policy_factory.py:10-16.
('SB3' is RL glue and out of scope)  Assigning these entries into the reference's own
`policy_factory` dict makes CrowdSimPlus humans run on the CUDA kernels."""
from .linear import Linear
from .orca import ORCA
from .orca_plus import ORCAPlus
from .social_force import SFM


def none_policy():
    return None


policy_factory = dict()
policy_factory['none'] = none_policy
policy_factory['linear'] = Linear
policy_factory['orca'] = ORCA
policy_factory['orca_plus'] = ORCAPlus
policy_factory['sfm'] = SFM
