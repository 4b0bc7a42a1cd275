"""rodeo_b200 -- B200-native drop-in for the batched probabilistic-ODE filtering path of mlysy/rodeo.

Mirrors the reference's public names (reference src/rodeo/__init__.py:1-6):
``solve_mv``, ``solve_sim``, ``interrogate``, ``prior``, ``inference``, ``utils``.
"""
__version__ = "0.1.0"

from . import interrogate, prior, utils, models, inference, kalmantv  # noqa: F401
from .prior import ibm_init  # noqa: F401
from .solve import solve_sim, solve_mv, solve_sim_loglik  # noqa: F401
