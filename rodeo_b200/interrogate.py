"""The reference's interrogation plug-ins (src/rodeo/interrogate.py), as handles to device code.

``solve_mv`` & co. recognise these function objects by identity -- also through
``functools.partial(interrogate_chkrebtii, kalman_type="standard")``, the spelling the reference's docs require
(docs/examples/parameter.md:331) -- and select the matching kernel instantiation.  They keep the reference's
signature but cannot be called on the host; an unknown callable raises ``NotImplementedError``.
"""
import functools

from . import _lib


def _device_only(name):
    raise NotImplementedError(f"{name} runs inside the CUDA kernels; pass it as `interrogate=` to a solver")


def interrogate_chkrebtii(key, ode_fun, ode_weight, t, mean_state_pred, var_state_pred, kalman_type, **params):
    """Chkrebtii et al (2016) -- reference src/rodeo/interrogate.py:13-47."""
    _device_only("interrogate_chkrebtii")


def interrogate_schober(key, ode_fun, ode_weight, t, mean_state_pred, var_state_pred, **params):
    """Schober et al (2019) -- reference src/rodeo/interrogate.py:50-62."""
    _device_only("interrogate_schober")


def interrogate_kramer(key, ode_fun, ode_weight, t, mean_state_pred, var_state_pred, **params):
    """Kramer et al (2021), first-order, block-diagonal Jacobian -- reference src/rodeo/interrogate.py:65-84."""
    _device_only("interrogate_kramer")


def interrogate_rodeo(key, ode_fun, ode_weight, t, mean_state_pred, var_state_pred, **params):
    """rodeo interrogation -- reference src/rodeo/interrogate.py:87-115."""
    _device_only("interrogate_rodeo")


_IDS = {
    interrogate_kramer: _lib.INTERROGATE_KRAMER,
    interrogate_chkrebtii: _lib.INTERROGATE_CHKREBTII,
    interrogate_schober: _lib.INTERROGATE_SCHOBER,
    interrogate_rodeo: _lib.INTERROGATE_RODEO,
}


for _fn, _id in _IDS.items():
    _fn.rodeo_id = _id          # the kernel enum, also reachable from a jax.ffi binding


def resolve(interrogate, kalman_type="standard"):
    """Map an interrogation callable to its kernel enum."""
    fn, kw = interrogate, {}
    while isinstance(fn, functools.partial):
        kw = {**fn.keywords, **kw}
        fn = fn.func
    if fn not in _IDS:
        raise NotImplementedError(
            "interrogate must be one of rodeo_b200.interrogate.interrogate_{kramer,chkrebtii,schober,rodeo} "
            "(optionally wrapped in functools.partial); custom interrogations are not compiled")
    if fn is interrogate_chkrebtii:
        # the reference never passes kalman_type itself (solve.py:70-78): it must be partial'd in
        if "kalman_type" not in kw:
            raise TypeError("interrogate_chkrebtii() missing 1 required positional argument: 'kalman_type' "
                            "(use functools.partial(interrogate_chkrebtii, kalman_type=...))")
        if kw["kalman_type"] != kalman_type:
            raise NotImplementedError("interrogate_chkrebtii kalman_type differs from the solver's kalman_type")
    return _IDS[fn]
