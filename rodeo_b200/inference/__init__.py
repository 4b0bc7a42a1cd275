"""Likelihood layers with the reference's names (src/rodeo/inference/__init__.py:1-3).

``rodeo_b200.inference.dalton`` is the function, as in the reference; the data-adaptive solvers of the reference's
``rodeo.inference.dalton`` *module* (``from rodeo.inference.dalton import solve_mv``) are reachable as
``rodeo_b200.inference.dalton_solve_mv`` / ``dalton_solve_sim`` / ``fenrir_solve_mv`` and under the module paths
``rodeo_b200.inference.dalton_module`` / ``fenrir_module``.
"""
from . import dalton as dalton_module
from .basic import basic
from . import fenrir as fenrir_module
from .fenrir import fenrir, solve_mv as fenrir_solve_mv
from .dalton import dalton, solve_mv as dalton_solve_mv, solve_sim as dalton_solve_sim
from .magi import magi_logdens
