from .policy import Policy
from .orca import ORCA
from .orca_plus import ORCAPlus
from .social_force import SFM
from .linear import Linear
from .policy_factory import policy_factory

__all__ = ["Policy", "ORCA", "ORCAPlus", "SFM", "Linear", "policy_factory"]
