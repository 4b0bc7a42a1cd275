r"""Batched square-root Kalman primitives on the GPU, with the names and keyword arguments of the reference's
``rodeo.kalmantv.square_root`` (src/rodeo/kalmantv/square_root.py:30-385).

Every variance argument and result is a lower-triangular factor ``L`` with ``var = L L^T`` (``forecast`` returns the
covariance itself, as the reference does, square_root.py:343-344).  The factors are built by Householder QR of the
stacked square roots (``rodeo.utils.add_sqrt``, src/rodeo/utils.py:10-24) and triangular solves, like the reference; a
QR leaves the signs of its triangular factor free, so only ``L L^T`` is comparable across implementations -- which is
also all that the reference's own tests compare (tests/test_square_root.py:11-16).  Arbitrary leading batch axes;
float64; CUDA tensors out.
"""
import torch

from .. import _host, _lib
from .standard import _batch, _flat, _prep, _stream


def predict(mean_state_past, var_state_past, mean_state, wgt_state, var_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:30-60 -> (mean_state_pred, var_state_pred [factor])"""
    m, L, c, Q, Rh = _prep(mean_state_past, var_state_past, mean_state, wgt_state, var_state)
    p = m.shape[-1]
    lead = torch.broadcast_shapes(_batch(m, 1), _batch(L, 2), _batch(c, 1), _batch(Q, 2), _batch(Rh, 2))
    m, c = _flat(m, lead, (p,)), _flat(c, lead, (p,))
    L, Q, Rh = (_flat(t, lead, (p, p)) for t in (L, Q, Rh))
    mo, Lo = torch.empty_like(m), torch.empty_like(L)
    rc = _lib.load().rodeo_b200_sqrt_predict_f64(m.shape[0], p, *map(_host.ptr, (m, L, c, Q, Rh, mo, Lo)), _stream())
    _lib.check(rc, "kalmantv.square_root.predict")
    return mo.reshape(*lead, p), Lo.reshape(*lead, p, p)


def _update_forecast(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, want_update, want_fore):
    m, L, d, W, Vh = _prep(mean_state_pred, var_state_pred, mean_meas, wgt_meas, var_meas)
    nm, p = W.shape[-2], W.shape[-1]
    x = _host.to_dev(x_meas) if x_meas is not None else torch.zeros_like(d)
    lead = torch.broadcast_shapes(_batch(m, 1), _batch(L, 2), _batch(x, 1), _batch(d, 1), _batch(W, 2), _batch(Vh, 2))
    m, x, d = _flat(m, lead, (p,)), _flat(x, lead, (nm,)), _flat(d, lead, (nm,))
    L, W, Vh = _flat(L, lead, (p, p)), _flat(W, lead, (nm, p)), _flat(Vh, lead, (nm, nm))
    mf = torch.empty_like(m) if want_update else None
    Lf = torch.empty_like(L) if want_update else None
    mz = torch.empty_like(d) if want_fore else None
    Sz = torch.empty_like(Vh) if want_fore else None
    rc = _lib.load().rodeo_b200_sqrt_update_f64(m.shape[0], p, nm, *map(_host.ptr, (m, L, x, d, W, Vh, mf, Lf, mz, Sz)),
                                                _stream())
    _lib.check(rc, "kalmantv.square_root.update")
    out = []
    if want_update:
        out += [mf.reshape(*lead, p), Lf.reshape(*lead, p, p)]
    if want_fore:
        out += [mz.reshape(*lead, nm), Sz.reshape(*lead, nm, nm)]
    return tuple(out)


def update(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:63-103 -> (mean_state_filt, var_state_filt [factor])"""
    return _update_forecast(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, True, False)


def forecast(mean_state_pred, var_state_pred, mean_meas, wgt_meas, var_meas, *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:317-345 -> (mean_fore, var_fore [a covariance])"""
    return _update_forecast(mean_state_pred, var_state_pred, None, mean_meas, wgt_meas, var_meas, False, True)


def filter(mean_state_past, var_state_past, mean_state, wgt_state, var_state, x_meas, mean_meas, wgt_meas, var_meas,
           *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:106-157 -> (mean_pred, var_pred, mean_filt, var_filt)"""
    mp, Lp = predict(mean_state_past, var_state_past, mean_state, wgt_state, var_state)
    mf, Lf = update(mp, Lp, x_meas, mean_meas, wgt_meas, var_meas)
    return mp, Lp, mf, Lf


def _smooth(mode, x_next, var_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state,
            var_state):
    mf, Lf, mp, Lp, Q, Rh = _prep(mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state, var_state)
    p = mf.shape[-1]
    xn = _host.to_dev(x_next) if x_next is not None else torch.zeros_like(mf)
    Ln = _host.to_dev(var_next) if var_next is not None else torch.zeros_like(Lf)
    lead = torch.broadcast_shapes(_batch(mf, 1), _batch(Lf, 2), _batch(mp, 1), _batch(Lp, 2), _batch(Q, 2), _batch(Rh, 2),
                                  _batch(xn, 1), _batch(Ln, 2))
    mf, mp, xn = (_flat(t, lead, (p,)) for t in (mf, mp, xn))
    Lf, Lp, Q, Rh, Ln = (_flat(t, lead, (p, p)) for t in (Lf, Lp, Q, Rh, Ln))
    om, ov, ow = torch.empty_like(mf), torch.empty_like(Lf), torch.empty_like(Lf)
    rc = _lib.load().rodeo_b200_sqrt_smooth_f64(mf.shape[0], p, mode,
                                                *map(_host.ptr, (xn, Ln, mf, Lf, mp, Lp, Q, Rh, om, ov, ow)), _stream())
    _lib.check(rc, "kalmantv.square_root.smooth")
    return om.reshape(*lead, p), ov.reshape(*lead, p, p), ow.reshape(*lead, p, p)


def smooth_mv(mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred,
              wgt_state, var_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:178-222 -> (mean_state_smooth, var_state_smooth [factor])"""
    m, L, _ = _smooth(0, mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred,
                      var_state_pred, wgt_state, var_state)
    return m, L


def smooth_sim(x_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state, var_state,
               *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:225-262 -> (mean_state_sim, var_state_sim [factor])"""
    m, L, _ = _smooth(1, x_state_next, None, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state,
                      var_state)
    return m, L


def smooth(x_state_next, mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred,
           var_state_pred, wgt_state, var_state, *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:265-315 -> (mean_sim, var_sim, mean_smooth, var_smooth)"""
    ms, Ls = smooth_sim(x_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state,
                        var_state)
    mm, Lm = smooth_mv(mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred,
                       var_state_pred, wgt_state, var_state)
    return ms, Ls, mm, Lm


def smooth_cond(mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state, var_state,
                *args, **kwargs):
    """reference src/rodeo/kalmantv/square_root.py:348-385 -> (wgt_state_cond, mean_state_cond, var_state_cond [factor])"""
    b, C, A = _smooth(2, None, None, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state,
                      var_state)
    return A, b, C
