"""Host-side helpers with the reference's names (src/rodeo/utils.py)."""
import ctypes

import numpy as np
import torch

from . import _host, _lib, models as _models


def first_order_pad(ode_fun, n_vars, n_deriv):
    r"""W and the initial-value helper for a first-order ODE (reference src/rodeo/utils.py:80-102).

    Returns ``(W, ode_init)``: ``W`` is the ``(n_vars, 1, n_deriv)`` NumPy matrix with ``W[:, :, 1] = 1``;
    ``ode_init(x0, t, theta=...)`` returns ``[x0, f(x0, t, theta), 0, ...]`` evaluated **on the device with the
    same functor the solver uses**, for ``x0`` of shape ``(n_vars,)`` or ``(B, n_vars)``.
    """
    model = _models.resolve(ode_fun)
    if (model.n_block, model.n_bstate) != (n_vars, n_deriv):
        raise ValueError(f"{model.name} is compiled for n_vars={model.n_block}, n_deriv={model.n_bstate}")

    def ode_init(x0, t, **params):
        dt = _host.real_dtype(params["theta"], x0)
        theta = _host.to_dev(params["theta"], dt)
        x0d = _host.to_dev(x0, dt)
        batched = theta.ndim == 2 or x0d.ndim == 2
        theta = theta[None] if theta.ndim == 1 else theta
        x0d = x0d[None] if x0d.ndim == 1 else x0d
        B = max(theta.shape[0], x0d.shape[0])
        theta = theta.expand(B, theta.shape[1]).contiguous()
        x0d = x0d.expand(B, n_vars).contiguous()
        c = _lib.RodeoProblem()
        c.B, c.n_steps, c.n_block, c.n_bstate, c.n_bmeas = B, 1, n_vars, n_deriv, model.n_bmeas
        c.n_theta, c.model_id, c.user_wcol = theta.shape[1], model.model_id, model.wcol
        X0 = torch.empty((B, n_vars, n_deriv), dtype=dt, device=_host.device())
        lib = _lib.load()
        fn = lib.rodeo_b200_ode_init_pad_f32 if dt == torch.float32 else lib.rodeo_b200_ode_init_pad_f64
        rc = fn(ctypes.byref(c), float(t), _host.ptr(theta), _host.ptr(x0d),
                                             _host.ptr(X0), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "ode_init")
        return X0 if batched else X0[0]

    W = np.zeros((n_vars, 1, n_deriv))
    W[:, :, 1] = 1.0
    return W, ode_init


def multivariate_normal_logpdf(x, mean, cov):
    """Eigendecomposition log-pdf with the reference's absolute 1e-8 eigenvalue cut-off (src/rodeo/utils.py:60-78),
    batched over leading axes; dimension 1..3; float64 CUDA tensor out."""
    x, mean, cov = _host.to_dev(x), _host.to_dev(mean), _host.to_dev(cov)
    n = cov.shape[-1]
    lead = torch.broadcast_shapes(x.shape[:-1], mean.shape[:-1], cov.shape[:-2])
    xf = x.expand(*lead, n).reshape(-1, n).contiguous()
    mf = mean.expand(*lead, n).reshape(-1, n).contiguous()
    cf = cov.expand(*lead, n, n).reshape(-1, n, n).contiguous()
    out = torch.empty((xf.shape[0],), dtype=torch.float64, device=xf.device)
    rc = _lib.load().rodeo_b200_mvn_logpdf_f64(xf.shape[0], n, _host.ptr(xf), _host.ptr(mf), _host.ptr(cf),
                                               _host.ptr(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "multivariate_normal_logpdf")
    return out.reshape(lead)
