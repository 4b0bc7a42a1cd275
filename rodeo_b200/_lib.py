"""ctypes binding of the C ABI declared in include/rodeo_b200.h.

The shared library is built in-tree (``make`` / ``__graft_entry__.build()``) as ``rodeo_b200/librodeo_b200.so``.
There is no CPU fallback: if the library is missing, or a compute entry point is called without a CUDA device,
the call fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RODEO_B200_LIB") or os.path.join(_HERE, "librodeo_b200.so")   # env: tuning builds

c_double_p = ctypes.c_void_p
c_int32_p = ctypes.c_void_p


class RodeoProblem(ctypes.Structure):
    """Mirror of ``struct RodeoProblem`` (include/rodeo_b200.h)."""
    _fields_ = [
        ("B", ctypes.c_int64),
        ("particle_offset", ctypes.c_int64),
        ("n_steps", ctypes.c_int32),
        ("n_block", ctypes.c_int32),
        ("n_bstate", ctypes.c_int32),
        ("n_bmeas", ctypes.c_int32),
        ("n_theta", ctypes.c_int32),
        ("model_id", ctypes.c_int32),
        ("interrogate", ctypes.c_int32),
        ("kalman_type", ctypes.c_int32),
        ("n_obs", ctypes.c_int32),
        ("n_bobs", ctypes.c_int32),
        ("key", ctypes.c_uint32 * 2),
        ("t_min", ctypes.c_double),
        ("t_max", ctypes.c_double),
        ("user_wcol", ctypes.c_int32),
        ("prior_batched", ctypes.c_int32),
        ("prior_var_scale", ctypes.c_void_p),
    ]


# enums of include/rodeo_b200.h
INTERROGATE_KRAMER, INTERROGATE_CHKREBTII, INTERROGATE_SCHOBER, INTERROGATE_RODEO = 0, 1, 2, 3
KALMAN_STANDARD, KALMAN_SQUARE_ROOT = 0, 1
OP_SOLVE_MV, OP_SOLVE_SIM, OP_DALTON, OP_FENRIR, OP_BASIC_GATHER, OP_ODE_INIT_PAD = range(6)
ERR_NAMES = {1: "UNSUPPORTED", 2: "INVALID", 3: "WORKSPACE", 4: "CUDA", 5: "NVRTC"}

_P = ctypes.POINTER(RodeoProblem)
_vp, _sz, _i, _d = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double

# name -> (restype, argtypes); every symbol include/rodeo_b200.h declares
SIGNATURES = {
    "rodeo_b200_workspace_bytes": (_sz, [_i, _P, _i]),
    "rodeo_b200_last_error": (ctypes.c_char_p, []),
    "rodeo_b200_abi_version": (_i, []),
    "rodeo_b200_problem_sizeof": (_sz, []),
    "rodeo_b200_launch_count": (ctypes.c_int64, []),
    "rodeo_b200_schedule_builds": (ctypes.c_int64, []),
    "rodeo_b200_schedule_clear": (None, []),
    "rodeo_b200_solve_mv_f64": (_i, [_P] + [_vp] * 8 + [_vp, _sz, _vp]),
    "rodeo_b200_solve_sim_f64": (_i, [_P] + [_vp] * 8 + [_vp, _sz, _vp]),
    "rodeo_b200_dalton_f64": (_i, [_P] + [_vp] * 11 + [_vp, _sz, _vp]),
    "rodeo_b200_fenrir_f64": (_i, [_P] + [_vp] * 11 + [_vp, _sz, _vp]),
    "rodeo_b200_basic_gather_f64": (_i, [_P, _vp, _vp, _vp, _vp]),
    "rodeo_b200_gauss_obs_loglik_f64": (_i, [_P, _vp, _vp, _vp, _d, _vp, _vp]),
    "rodeo_b200_solve_sim_loglik_f64": (_i, [_P] + [_vp] * 9 + [_d, _vp, _vp] + [_vp, _sz, _vp]),
    "rodeo_b200_rwmh_propose_f64": (_i, [ctypes.c_int64, _i, _vp, _vp, _vp, _vp, ctypes.c_int64, _vp, _vp]),
    "rodeo_b200_rwmh_accept_f64": (_i, [ctypes.c_int64, _i, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int64, _vp, _vp,
                                        _vp]),
    "rodeo_b200_ode_init_pad_f64": (_i, [_P, _d, _vp, _vp, _vp, _vp]),
    "rodeo_b200_solve_mv_f32": (_i, [_P] + [_vp] * 8 + [_vp, _sz, _vp]),
    "rodeo_b200_solve_sim_f32": (_i, [_P] + [_vp] * 8 + [_vp, _sz, _vp]),
    "rodeo_b200_dalton_f32": (_i, [_P] + [_vp] * 11 + [_vp, _sz, _vp]),
    "rodeo_b200_fenrir_f32": (_i, [_P] + [_vp] * 11 + [_vp, _sz, _vp]),
    "rodeo_b200_basic_gather_f32": (_i, [_P, _vp, _vp, _vp, _vp]),
    "rodeo_b200_ode_init_pad_f32": (_i, [_P, ctypes.c_float, _vp, _vp, _vp, _vp]),
    "rodeo_b200_dalton_solve_workspace_bytes_f64": (_sz, [_i, _P]),
    "rodeo_b200_dalton_solve_mv_f64": (_i, [_P] + [_vp] * 12 + [_vp, _sz, _vp]),
    "rodeo_b200_dalton_solve_sim_f64": (_i, [_P] + [_vp] * 12 + [_vp, _sz, _vp]),
    "rodeo_b200_solve_mv_sqrt_workspace_bytes": (_sz, [_P]),
    "rodeo_b200_solve_mv_sqrt_f64": (_i, [_P] + [_vp] * 8 + [_vp, _sz, _vp]),
    "rodeo_b200_solve_mv_sqrt_f32": (_i, [_P] + [_vp] * 8 + [_vp, _sz, _vp]),
    "rodeo_b200_fenrir_solve_mv_workspace_bytes": (_sz, [_P]),
    "rodeo_b200_fenrir_solve_mv_f64": (_i, [_P] + [_vp] * 12 + [_vp, _sz, _vp]),
    "rodeo_b200_ktv_predict_f64": (_i, [ctypes.c_int64, _i] + [_vp] * 7 + [_vp]),
    "rodeo_b200_ktv_update_f64": (_i, [ctypes.c_int64, _i, _i] + [_vp] * 10 + [_vp]),
    "rodeo_b200_ktv_smooth_f64": (_i, [ctypes.c_int64, _i, _i] + [_vp] * 10 + [_vp]),
    "rodeo_b200_mvn_logpdf_f64": (_i, [ctypes.c_int64, _i] + [_vp] * 4 + [_vp]),
    "rodeo_b200_sqrt_predict_f64": (_i, [ctypes.c_int64, _i] + [_vp] * 7 + [_vp]),
    "rodeo_b200_sqrt_update_f64": (_i, [ctypes.c_int64, _i, _i] + [_vp] * 10 + [_vp]),
    "rodeo_b200_sqrt_smooth_f64": (_i, [ctypes.c_int64, _i, _i] + [_vp] * 11 + [_vp]),
    "rodeo_b200_psd_factor_f64": (_i, [ctypes.c_int64, _i, _vp, _vp, _vp]),
    "rodeo_b200_magi_logdens_f64": (_i, [ctypes.c_int64, _i, _i, _i, _i] + [_vp] * 4 + [_vp]),
    "rodeo_b200_dalton_f64_host": (_i, [_P] + [_vp] * 10),
    "rodeo_b200_solve_mv_f64_host": (_i, [_P] + [_vp] * 7),
    "rodeo_b200_host_arena_release": (None, []),
    "rodeo_b200_fp64_peak_probe": (_i, [_i, _vp]),
    "rodeo_b200_register_model_nvrtc": (_i, [ctypes.c_char_p, ctypes.c_char_p, _i, _i, _i, _i, _vp]),
    "rodeo_b200_peer_region_bytes": (_sz, [ctypes.c_longlong, _i]),
    "rodeo_b200_peer_alloc": (_i, [_sz, ctypes.POINTER(_vp), _vp]),
    "rodeo_b200_peer_open": (_i, [_vp, ctypes.POINTER(_vp)]),
    "rodeo_b200_peer_close": (_i, [_vp]),
    "rodeo_b200_peer_free": (_i, [_vp]),
    "rodeo_b200_peer_allgather_f64": (_i, [_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _vp,
                                           ctypes.c_uint, _vp, ctypes.c_ulonglong, _vp, _vp]),
}

_lib = None


class RodeoError(RuntimeError):
    pass


def load():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RodeoError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "rodeo_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.rodeo_b200_abi_version() != 1:
        raise RodeoError("librodeo_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().rodeo_b200_last_error().decode(errors="replace")
        exc = NotImplementedError if rc == 1 else (ValueError if rc == 2 else RodeoError)
        raise exc(f"{what} failed [{ERR_NAMES.get(rc, rc)}]: {msg}")
