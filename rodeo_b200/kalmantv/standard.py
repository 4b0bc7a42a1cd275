r"""Batched covariance-form Kalman primitives on the GPU, with the names and keyword arguments of the reference's
``rodeo.kalmantv.standard`` (src/rodeo/kalmantv/standard.py).

Every argument may carry arbitrary leading batch axes (the reference is un-batched and ``jax.vmap``-ed); the kernels run
one thread per problem and call the very ``__device__`` functions that the fused solver kernels inline, which is what
lets the reference's own known-answer tests of these primitives be restated against the CUDA code.  float64, CUDA
tensors out.
"""
import ctypes

import torch

from .. import _host, _lib


def _prep(*arrs):
    ts = [_host.to_dev(a) for a in arrs]
    return ts


def _batch(t, nd):
    """split leading batch axes from the trailing `nd` problem axes"""
    return t.shape[:t.dim() - nd]


def _flat(t, lead, tail):
    return t.expand(*lead, *tail).reshape(-1, *tail).contiguous()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def predict(mean_state_past, var_state_past, mean_state, wgt_state, var_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:31-60 -> (mean_state_pred, var_state_pred)"""
    m, S, c, Q, R = _prep(mean_state_past, var_state_past, mean_state, wgt_state, var_state)
    p = m.shape[-1]
    lead = torch.broadcast_shapes(_batch(m, 1), _batch(S, 2), _batch(c, 1), _batch(Q, 2), _batch(R, 2))
    m, c = _flat(m, lead, (p,)), _flat(c, lead, (p,))
    S, Q, R = (_flat(t, lead, (p, p)) for t in (S, Q, R))
    B = m.shape[0]
    mo, So = torch.empty_like(m), torch.empty_like(S)
    rc = _lib.load().rodeo_b200_ktv_predict_f64(B, p, *map(_host.ptr, (m, S, c, Q, R, mo, So)), _stream())
    _lib.check(rc, "kalmantv.predict")
    return mo.reshape(*lead, p), So.reshape(*lead, p, p)


def _update_forecast(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, want_update, want_fore):
    m, S, d, W, V = _prep(mean_state_pred, var_state_pred, mean_meas, wgt_meas, var_meas)
    nm, p = W.shape[-2], W.shape[-1]
    x = _host.to_dev(x_meas) if x_meas is not None else torch.zeros_like(d)
    lead = torch.broadcast_shapes(_batch(m, 1), _batch(S, 2), _batch(x, 1), _batch(d, 1), _batch(W, 2), _batch(V, 2))
    m, x, d = _flat(m, lead, (p,)), _flat(x, lead, (nm,)), _flat(d, lead, (nm,))
    S, W, V = _flat(S, lead, (p, p)), _flat(W, lead, (nm, p)), _flat(V, lead, (nm, nm))
    B = m.shape[0]
    mf = torch.empty_like(m) if want_update else None
    Sf = torch.empty_like(S) if want_update else None
    mz = torch.empty_like(d) if want_fore else None
    Sz = torch.empty_like(V) if want_fore else None
    rc = _lib.load().rodeo_b200_ktv_update_f64(B, p, nm, *map(_host.ptr, (m, S, x, d, W, V, mf, Sf, mz, Sz)), _stream())
    _lib.check(rc, "kalmantv.update")
    out = []
    if want_update:
        out += [mf.reshape(*lead, p), Sf.reshape(*lead, p, p)]
    if want_fore:
        out += [mz.reshape(*lead, nm), Sz.reshape(*lead, nm, nm)]
    return tuple(out)


def update(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:63-103 -> (mean_state_filt, var_state_filt)"""
    return _update_forecast(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, True, False)


def forecast(mean_state_pred, var_state_pred, mean_meas, wgt_meas, var_meas, *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:308-336 -> (mean_fore, var_fore)"""
    return _update_forecast(mean_state_pred, var_state_pred, None, mean_meas, wgt_meas, var_meas, False, True)


def filter(mean_state_past, var_state_past, mean_state, wgt_state, var_state, x_meas, mean_meas, wgt_meas, var_meas,
           *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:106-157 -> (mean_pred, var_pred, mean_filt, var_filt)"""
    mp, Sp = predict(mean_state_past, var_state_past, mean_state, wgt_state, var_state)
    mf, Sf = update(mp, Sp, x_meas, mean_meas, wgt_meas, var_meas)
    return mp, Sp, mf, Sf


def _smooth(mode, x_next, var_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state):
    mf, Sf, mp, Sp, Q = _prep(mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state)
    p = mf.shape[-1]
    xn = _host.to_dev(x_next) if x_next is not None else torch.zeros_like(mf)
    Sn = _host.to_dev(var_next) if var_next is not None else torch.zeros_like(Sf)
    lead = torch.broadcast_shapes(_batch(mf, 1), _batch(Sf, 2), _batch(mp, 1), _batch(Sp, 2), _batch(Q, 2),
                                  _batch(xn, 1), _batch(Sn, 2))
    mf, mp, xn = (_flat(t, lead, (p,)) for t in (mf, mp, xn))
    Sf, Sp, Q, Sn = (_flat(t, lead, (p, p)) for t in (Sf, Sp, Q, Sn))
    B = mf.shape[0]
    om, ov, ow = torch.empty_like(mf), torch.empty_like(Sf), torch.empty_like(Sf)
    rc = _lib.load().rodeo_b200_ktv_smooth_f64(B, p, mode, *map(_host.ptr, (xn, Sn, mf, Sf, mp, Sp, Q, om, ov, ow)),
                                               _stream())
    _lib.check(rc, "kalmantv.smooth")
    return om.reshape(*lead, p), ov.reshape(*lead, p, p), ow.reshape(*lead, p, p)


def smooth_mv(mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred,
              wgt_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:180-217 -> (mean_state_smooth, var_state_smooth)"""
    m, v, _ = _smooth(0, mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred,
                      var_state_pred, wgt_state)
    return m, v


def smooth_sim(x_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state,
               *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:220-255 -> (mean_state_sim, var_state_sim)"""
    m, v, _ = _smooth(1, x_state_next, None, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred,
                      wgt_state)
    return m, v


def smooth(x_state_next, mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred,
           var_state_pred, wgt_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:258-305 -> (mean_sim, var_sim, mean_smooth, var_smooth)"""
    ms, vs = smooth_sim(x_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state)
    mm, vm = smooth_mv(mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred,
                       var_state_pred, wgt_state)
    return ms, vs, mm, vm


def smooth_cond(mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/standard.py:339-371 -> (wgt_state_cond, mean_state_cond, var_state_cond)"""
    b, C, A = _smooth(2, None, None, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state)
    return A, b, C


def psd_factor(var):
    """Lower-triangular ``A`` with ``A A^T = var`` for a positive semi-definite ``var`` (``[..., p, p]``): the factor
    the sampling kernels draw with, where the reference uses ``jax.random.multivariate_normal`` with its SVD / Cholesky
    factor (src/rodeo/solve.py:179; src/rodeo/interrogate.py:30-34)."""
    V, = _prep(var)
    p = V.shape[-1]
    lead = _batch(V, 2)
    Vf = _flat(V, lead, (p, p))
    A = torch.empty_like(Vf)
    rc = _lib.load().rodeo_b200_psd_factor_f64(Vf.shape[0], p, _host.ptr(Vf), _host.ptr(A), _stream())
    _lib.check(rc, "kalmantv.psd_factor")
    return A.reshape(*lead, p, p)
