"""Drop-in ``rodeo.inference.basic`` (reference src/rodeo/inference/basic.py:16-62) on the B200 kernels."""
import ctypes

import torch

from .. import _host, _lib
from ..solve import solve_mv


def basic(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
          obs_data=None, obs_times=None, obs_loglik=None, kalman_type="standard",
          prior_weight=None, prior_var=None, **params):
    r"""Basic log-likelihood: ``solve_mv`` then ``obs_loglik(obs_data, Xt[obs_ind], **params)``.

    ``obs_loglik`` is the user's own function, as in the reference; it receives the gathered solver means
    ``ode_data`` of shape ``([B,] n_obs, n_block, n_bstate)`` as a CUDA tensor (the gather runs on the device).
    Returns ``(obs_loglik(...), Xt)`` exactly like the reference (basic.py:62).
    """
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params)
    pb.set_obs(obs_data, obs_times)
    Xt, _ = solve_mv(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate,
                     prior_pars=prior_pars, prior_weight=prior_weight, prior_var=prior_var,
                     kalman_type=kalman_type, **params)      # the caller's own prior (may be per theta)
    Xb = Xt if pb.batched else Xt[None]
    ode_data = pb.empty(pb.B, pb.c.n_obs, pb.nb, pb.p)
    rc = pb.fn("basic_gather")(ctypes.byref(pb.c), _host.ptr(Xb), _host.ptr(pb.obs_ind),
                                            _host.ptr(ode_data), pb.stream())
    _lib.check(rc, "basic_gather")
    return obs_loglik(obs_data, pb.unbatch(ode_data), **params), Xt
