"""Drop-in ``rodeo.inference.dalton`` (reference src/rodeo/inference/dalton.py:39-235) on the B200 kernel."""
import ctypes

import torch

from .. import _host, _lib


def dalton(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
           obs_data=None, obs_times=None, obs_weight=None, obs_var=None, kalman_type="standard",
           prior_weight=None, prior_var=None, _z_interr=None, _particle_offset=0, **params):
    r"""DALTON marginal log-likelihood :math:`\log p(Y_{0:M} \mid Z_{1:N})` for Gaussian observations.

    Same arguments as the reference; ``theta`` / ``ode_init`` may carry a leading batch axis ``B``.
    Returns a float64 CUDA tensor of shape ``(B,)`` (a 0-d tensor for an un-batched call).
    """
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params, particle_offset=_particle_offset)
    pb.set_obs(obs_data, obs_times, obs_weight, obs_var)
    out = pb.empty(pb.B)
    zi = None if _z_interr is None else pb.dev(_z_interr)
    rc = pb.fn("dalton")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                      _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(pb.obs_ind),
                                      _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight), _host.ptr(pb.obs_var),
                                      _host.ptr(out), None, 0, pb.stream())
    _lib.check(rc, "dalton")
    return pb.unbatch(out)
