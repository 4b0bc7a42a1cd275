"""ctypes loader of the C/OpenMP oracle port (oracle/rodeo_oracle.c)  --  TEST / BASELINE INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librodeo_oracle.so")
MODEL_IDS = {"fitzhugh_nagumo": 0, "lorenz63": 1, "second_order_sin": 2, "hes1": 3, "seirah": 4}
INTERR_IDS = {"kramer": 0, "chkrebtii": 1, "schober": 2, "rodeo": 3}
_SO_LD = os.path.join(_HERE, "_build", "librodeo_oracle_ld.so")
_lib = None
_lib_ld = None


def build():
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.rodeo_oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt=np.float64):
    return np.ascontiguousarray(np.asarray(a, dtype=dt))


def max_threads():
    return int(load().rodeo_oracle_max_threads())


def load_ld():
    """the same C source compiled with `long double` arithmetic (x87, 64-bit mantissa)"""
    global _lib_ld
    if _lib_ld is None:
        if not os.path.exists(_SO_LD):
            build()
        _lib_ld = ctypes.CDLL(_SO_LD)
    return _lib_ld


def dalton_ld(*args, **kw):
    """dalton evaluated in extended precision and rounded once to float64: the 'exact' value the float64 noise floor
    of the log-likelihood is measured against (same recursion, ~2000x finer rounding)."""
    return dalton(*args, _lib_override=load_ld(), **kw)


def dalton(model, interr, W, X0, t_min, t_max, n_steps, Q, R, theta, obs_data, obs_ind, obs_weight, obs_var,
           n_threads=0, _lib_override=None):
    lib = _lib_override or load()
    W, X0, Q, R, theta = _c(W), _c(X0), _c(Q), _c(R), _c(theta)
    obs_data, obs_weight, obs_var, obs_ind = _c(obs_data), _c(obs_weight), _c(obs_var), _c(obs_ind, np.int32)
    B, nb, p = X0.shape
    out = np.empty(B)
    rc = lib.rodeo_oracle_dalton(
        ctypes.c_int(MODEL_IDS[model]), ctypes.c_int(INTERR_IDS[interr]), ctypes.c_long(B), ctypes.c_int(n_steps),
        ctypes.c_int(nb), ctypes.c_int(p), ctypes.c_int(theta.shape[1]), ctypes.c_double(t_min),
        ctypes.c_double(t_max), _p(W), _p(Q), _p(R), _p(X0), _p(theta), ctypes.c_int(len(obs_ind)), _p(obs_ind),
        _p(obs_data), _p(obs_weight), _p(obs_var), _p(out), ctypes.c_int(n_threads))
    if rc:
        raise ValueError("unsupported configuration for the C oracle port")
    return out


def solve_mv(model, interr, W, X0, t_min, t_max, n_steps, Q, R, theta, n_threads=0):
    lib = load()
    W, X0, Q, R, theta = _c(W), _c(X0), _c(Q), _c(R), _c(theta)
    B, nb, p = X0.shape
    mean = np.empty((B, n_steps + 1, nb, p)); var = np.empty((B, n_steps + 1, nb, p, p))
    rc = lib.rodeo_oracle_solve_mv(
        ctypes.c_int(MODEL_IDS[model]), ctypes.c_int(INTERR_IDS[interr]), ctypes.c_long(B), ctypes.c_int(n_steps),
        ctypes.c_int(nb), ctypes.c_int(p), ctypes.c_int(theta.shape[1]), ctypes.c_double(t_min),
        ctypes.c_double(t_max), _p(W), _p(Q), _p(R), _p(X0), _p(theta), _p(mean), _p(var), ctypes.c_int(n_threads))
    if rc:
        raise ValueError("unsupported configuration for the C oracle port")
    return mean, var
