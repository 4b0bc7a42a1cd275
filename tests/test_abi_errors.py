"""CPU: error behaviour of the C ABI that is decided before any CUDA call (return code + rodeo_b200_last_error()).

The reference validates almost nothing (only `NotImplementedError` for an unknown kalman_type, src/rodeo/solve.py:236-241);
the boundary adds the checks a pointer-based ABI needs and reports them through return codes, never by crashing.
"""
import ctypes

import numpy as np

from rodeo_b200 import _lib


def _problem(**kw):
    c = _lib.RodeoProblem()
    c.B, c.n_steps, c.n_block, c.n_bstate, c.n_bmeas, c.n_theta = 4, 10, 2, 3, 1, 3
    c.model_id, c.interrogate, c.kalman_type, c.n_obs, c.n_bobs = 0, 0, 0, 2, 1
    c.t_min, c.t_max, c.user_wcol = 0.0, 1.0, 1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _err(lib):
    return lib.rodeo_b200_last_error().decode()


def _call_solve_mv(lib, c, ws=None, ws_bytes=0):
    W = np.zeros((2, 1, 3)); W[:, :, 1] = 1
    Q = np.tile(np.eye(3), (2, 1, 1)); R = np.tile(np.eye(3), (2, 1, 1))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    return lib.rodeo_b200_solve_mv_f64(ctypes.byref(c), p(W), p(Q), p(R), None, None, None, None, None, ws, ws_bytes, None)


def test_unknown_kalman_type_is_unsupported():
    lib = _lib.load()
    assert _call_solve_mv(lib, _problem(kalman_type=7)) == 1           # RODEO_ERR_UNSUPPORTED
    assert "kalman_type" in _err(lib)


def test_bad_sizes_are_invalid():
    lib = _lib.load()
    assert _call_solve_mv(lib, _problem(n_steps=0)) == 2               # RODEO_ERR_INVALID
    assert _call_solve_mv(lib, _problem(B=-1)) == 2


def test_workspace_too_small_is_reported_not_crashed():
    lib = _lib.load()
    c = _problem()
    need = lib.rodeo_b200_workspace_bytes(_lib.OP_SOLVE_MV, ctypes.byref(c), 8)
    assert need > 0
    assert _call_solve_mv(lib, c, None, 0) == 3                        # RODEO_ERR_WORKSPACE
    assert str(need) in _err(lib)


def test_model_dimension_mismatch_and_unknown_model():
    lib = _lib.load()
    c = _problem(n_block=3)                                            # FitzHugh-Nagumo has 2 blocks
    ws = ctypes.create_string_buffer(1 << 20)
    W = np.zeros((3, 1, 3)); Q = np.tile(np.eye(3), (3, 1, 1))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.rodeo_b200_solve_mv_f64(ctypes.byref(c), p(W), p(Q), p(Q), None, None, None, None, None,
                                     ctypes.cast(ws, ctypes.c_void_p), 1 << 20, None)
    assert rc == 2 and "FitzHughNagumo expects" in _err(lib)
    c = _problem(model_id=77)
    rc = _call_solve_mv(lib, c, ctypes.cast(ws, ctypes.c_void_p), 1 << 20)
    assert rc == 1 and "not compiled" in _err(lib)


def test_null_problem():
    lib = _lib.load()
    assert lib.rodeo_b200_solve_mv_f64(None, None, None, None, None, None, None, None, None, None, 0, None) == 2
    assert lib.rodeo_b200_workspace_bytes(_lib.OP_SOLVE_MV, None, 8) == 0


def test_per_theta_prior_is_refused_where_it_is_not_compiled():
    """RodeoProblem.prior_batched: float64 solve_mv / solve_sim / dalton / fenrir only; everything else must say so"""
    lib = _lib.load()
    ws = ctypes.create_string_buffer(1 << 20)
    W = np.zeros((2, 1, 3)); W[:, :, 1] = 1
    Q = np.tile(np.eye(3), (2, 1, 1))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    # user (NVRTC) models: refused by the common checks
    c = _problem(prior_batched=1, model_id=1000)
    assert _call_solve_mv(lib, c, ctypes.cast(ws, ctypes.c_void_p), 1 << 20) == 1 and "per-theta prior" in _err(lib)
    # float32 entry points carry no QK_DENSE_BATCH instantiation
    c = _problem(prior_batched=1)
    Wf, Qf = W.astype(np.float32), Q.astype(np.float32)
    rc = lib.rodeo_b200_solve_mv_f32(ctypes.byref(c), p(Wf), p(Qf), p(Qf), None, None, None, None, None,
                                     ctypes.cast(ws, ctypes.c_void_p), 1 << 20, None)
    assert rc == 1 and "per-theta prior" in _err(lib)
    # the host-buffer wrappers take a shared prior
    rc = lib.rodeo_b200_solve_mv_f64_host(ctypes.byref(c), p(W), p(Q), p(Q), None, None, None, None)
    assert rc == 1 and "shared prior" in _err(lib)


def test_two_measurement_rows_model_is_compiled_for_the_float64_solvers_only():
    lib = _lib.load()
    ws = ctypes.create_string_buffer(1 << 20)
    c = _problem(model_id=5, n_block=1, n_bstate=6, n_bmeas=2)
    W = np.zeros((1, 2, 6), np.float32); Q = np.tile(np.eye(6, dtype=np.float32), (1, 1, 1))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.rodeo_b200_solve_mv_f32(ctypes.byref(c), p(W), p(Q), p(Q), None, None, None, None, None,
                                     ctypes.cast(ws, ctypes.c_void_p), 1 << 20, None)
    assert rc == 1 and "not compiled" in _err(lib)


def test_peer_gather_argument_checks():
    lib = _lib.load()
    assert lib.rodeo_b200_peer_region_bytes(1000, 8) >= 2 * 1000 * 8 + 2 * 8 * 4
    assert lib.rodeo_b200_peer_region_bytes(1000, 17) == 0 and lib.rodeo_b200_peer_region_bytes(-1, 2) == 0
    # NULL shard / regions, rank out of range, shard past the end, no epoch source: refused before any CUDA call
    assert lib.rodeo_b200_peer_allgather_f64(None, 4, 0, 8, 0, 2, None, 1, None, 0, None, None) == 2
    # (an empty shard may come with a NULL pointer: only a NULL shard of positive length is refused)
    dummy = ctypes.create_string_buffer(64)
    regs = (ctypes.c_void_p * 2)(ctypes.addressof(dummy), ctypes.addressof(dummy))
    p = ctypes.cast(dummy, ctypes.c_void_p)
    for args in ((p, 4, 0, 8, 2, 2, regs, 1, None, 0, p, None),          # rank >= world
                 (p, 4, 6, 8, 0, 2, regs, 1, None, 0, p, None),          # offset + n_local > n_total
                 (p, 4, 0, 8, 0, 2, regs, 0, None, 0, p, None)):         # epoch 0 without a device counter
        assert lib.rodeo_b200_peer_allgather_f64(*args[:6], ctypes.cast(args[6], ctypes.c_void_p), *args[7:]) == 2
        assert "peer_allgather" in _err(lib)
