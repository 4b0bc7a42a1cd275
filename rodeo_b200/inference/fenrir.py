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
