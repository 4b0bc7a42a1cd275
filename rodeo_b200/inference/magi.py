"""Drop-in ``rodeo.inference.magi_logdens`` (reference src/rodeo/inference/magi.py:6-99) on a B200 kernel."""
import ctypes

import torch

from .. import _host, _lib


def magi_logdens(ode_data_subset, ode_expand, n_active, prior_pars, kalman_type="standard", **params):
    r"""Log-density of the MAGI approximation, :math:`p(U_{0:N}, Z = 0 \mid \theta)`.

    Same arguments as the reference.  ``ode_expand(ode_data_subset, **params)`` is the user's own function, as in the
    reference; it is called once, on the host side, and must return the full solution process of shape
    ``([B,] n_steps + 1, n_block, n_bstate)`` (NumPy, or a torch tensor on any device -- nothing is traced or
    differentiated here, so any array code works).  A leading batch axis on its result batches the evaluation over
    trajectories / parameters.  Returns a float64 CUDA tensor of shape ``([B])``.
    """
    if kalman_type != "standard":
        raise NotImplementedError(f"kalman_type={kalman_type!r} (only \"standard\" is built for magi_logdens)")
    X = _host.to_dev(ode_expand(ode_data_subset, **params))
    batched = X.dim() == 4
    if X.dim() not in (3, 4):
        raise ValueError(f"ode_expand must return ([B,] n_steps+1, n_block, n_bstate); got shape {tuple(X.shape)}")
    Xb = X if batched else X[None]
    B, N1, nb, p = Xb.shape
    Q, R = (_host.to_host(a) for a in prior_pars)
    if Q.shape != (nb, p, p) or R.shape != (nb, p, p):
        raise ValueError(f"prior_pars must be two (n_block, n_bstate, n_bstate) arrays; got {Q.shape}, {R.shape}")
    out = torch.empty((B,), dtype=torch.float64, device=Xb.device)
    rc = _lib.load().rodeo_b200_magi_logdens_f64(B, N1 - 1, nb, p, int(n_active), _host.ptr(Q), _host.ptr(R),
                                                 _host.ptr(Xb), _host.ptr(out),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "magi_logdens")
    return out if batched else out[0]
