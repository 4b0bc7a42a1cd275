"""Kalman primitives with the reference's module layout (src/rodeo/kalmantv/)."""
from . import standard  # noqa: F401
