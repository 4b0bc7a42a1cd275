"""Drop-in ``rodeo.inference.fenrir`` (reference src/rodeo/inference/fenrir.py:261-328) on the B200 kernel."""
import ctypes

import torch

from .. import _host, _lib


def fenrir(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
           obs_data=None, obs_times=None, obs_weight=None, obs_var=None, kalman_type="standard",
           prior_weight=None, prior_var=None, _z_interr=None, **params):
    r"""Fenrir log-likelihood :math:`\log p(Y_{0:M} \mid Z_{1:N})` (Tronarp et al 2022).

    Same arguments as the reference; returns a float64 CUDA tensor ``(B,)`` (0-d if un-batched).
    """
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params)
    pb.set_obs(obs_data, obs_times, obs_weight, obs_var)
    if kalman_type != "standard":
        raise NotImplementedError('only kalman_type="standard" is built for the log-likelihood layers')
    out = pb.empty(pb.B)
    ws, n = pb.workspace(_lib.OP_FENRIR)
    zi = None if _z_interr is None else pb.dev(_z_interr)
    rc = pb.fn("fenrir")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                      _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(pb.obs_ind),
                                      _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight), _host.ptr(pb.obs_var),
                                      _host.ptr(out), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "fenrir")
    return pb.unbatch(out)


def solve_mv(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
             obs_data=None, obs_times=None, obs_weight=None, obs_var=None, kalman_type="standard",
             prior_weight=None, prior_var=None, _z_interr=None, **params):
    r"""Fenrir posterior mean / variance of :math:`p(X_{0:N} \mid Z_{1:N}, Y_{0:M})`
    (reference src/rodeo/inference/fenrir.py:404-457).  Same arguments as :func:`fenrir`."""
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params)
    pb.set_obs(obs_data, obs_times, obs_weight, obs_var)
    if pb.sfx != "f64" or kalman_type != "standard":
        raise NotImplementedError('fenrir.solve_mv is compiled for float64, kalman_type="standard" only')
    N = pb.n_steps
    mean, var = pb.empty(pb.B, N + 1, pb.nb, pb.p), pb.empty(pb.B, N + 1, pb.nb, pb.p, pb.p)
    n = pb.lib.rodeo_b200_fenrir_solve_mv_workspace_bytes(ctypes.byref(pb.c))
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=_host.device())
    zi = None if _z_interr is None else pb.dev(_z_interr)
    rc = pb.lib.rodeo_b200_fenrir_solve_mv_f64(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                               _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi),
                                               _host.ptr(pb.obs_ind), _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight),
                                               _host.ptr(pb.obs_var), _host.ptr(mean), _host.ptr(var), _host.ptr(ws), n,
                                               pb.stream())
    _lib.check(rc, "fenrir.solve_mv")
    return pb.unbatch(mean), pb.unbatch(var)
