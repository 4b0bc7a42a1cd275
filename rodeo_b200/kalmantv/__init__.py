"""Kalman primitives with the reference's module layout (src/rodeo/kalmantv/)."""
from . import standard, square_root  # noqa: F401
