"""Drop-in ``rodeo.inference.dalton`` (reference src/rodeo/inference/dalton.py:39-235) on the B200 kernel."""
import ctypes

import torch

from .. import _host, _lib


def dalton(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
           obs_data=None, obs_times=None, obs_weight=None, obs_var=None, kalman_type="standard",
           prior_weight=None, prior_var=None, _z_interr=None, _particle_offset=0, **params):
    r"""DALTON marginal log-likelihood :math:`\log p(Y_{0:M} \mid Z_{1:N})` for Gaussian observations.

    Same arguments as the reference; ``theta`` / ``ode_init`` may carry a leading batch axis ``B``.
    Returns a tensor of shape ``(B,)`` (a 0-d tensor for an un-batched call) where the inputs live: on the CUDA device
    when ``theta`` / ``ode_init`` / the observation arrays are CUDA tensors; on the host when all of them are host
    arrays (NumPy, as a user of the reference has them) -- that call goes through ``rodeo_b200_dalton_f64_host``, which
    cuts the batch in two and sends the second half while the kernel of the first one runs.
    """
    if kalman_type != "standard":
        raise NotImplementedError('only kalman_type="standard" is built for the log-likelihood layers')
    host = _z_interr is None and obs_weight is not None and \
        _host.on_host(params.get("theta"), ode_init, obs_data, obs_weight, obs_var) and \
        _host.real_dtype(params.get("theta"), ode_init) == torch.float64
    if host:
        hp = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                           prior_weight, prior_var, kalman_type, params, particle_offset=_particle_offset,
                           host_inputs=True)
        if hp.prior_batch is None:                   # (a general per-theta prior takes the device entry point)
            hp.set_obs(obs_data, obs_times, obs_weight, obs_var)
            out = hp.empty(hp.B)
            rc = hp.lib.rodeo_b200_dalton_f64_host(ctypes.byref(hp.c), _host.ptr(hp.W), _host.ptr(hp.Q), _host.ptr(hp.R),
                                                   _host.ptr(hp.x0), _host.ptr(hp.theta), _host.ptr(hp.obs_ind),
                                                   _host.ptr(hp.obs_data), _host.ptr(hp.obs_weight),
                                                   _host.ptr(hp.obs_var), _host.ptr(out))
            _lib.check(rc, "dalton (host buffers)")
            return hp.unbatch(out)
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


def _adaptive(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, obs_data, obs_times,
              obs_weight, obs_var, kalman_type, prior_weight, prior_var, params, particle_offset=0):
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params, particle_offset=particle_offset)
    pb.set_obs(obs_data, obs_times, obs_weight, obs_var)
    if pb.sfx != "f64" or kalman_type != "standard":
        raise NotImplementedError("the data-adaptive solvers are compiled for float64, kalman_type=\"standard\" only")
    return pb


def solve_mv(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
             obs_data=None, obs_times=None, obs_weight=None, obs_var=None, kalman_type="standard",
             prior_weight=None, prior_var=None, _z_interr=None, **params):
    r"""DALTON data-adaptive posterior mean / variance :math:`p(X_{0:N} \mid Y_{0:M}, Z_{1:N})`
    (reference src/rodeo/inference/dalton.py:374-460).  Same arguments as :func:`dalton`; returns ``(mean, var)``
    with the shapes of :func:`rodeo_b200.solve_mv`."""
    pb = _adaptive(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, obs_data,
                   obs_times, obs_weight, obs_var, kalman_type, prior_weight, prior_var, params)
    N = pb.n_steps
    mean, var = pb.empty(pb.B, N + 1, pb.nb, pb.p), pb.empty(pb.B, N + 1, pb.nb, pb.p, pb.p)
    n = getattr(pb.lib, f"rodeo_b200_dalton_solve_workspace_bytes_{pb.sfx}")(_lib.OP_SOLVE_MV, ctypes.byref(pb.c))
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=_host.device())
    zi = None if _z_interr is None else pb.dev(_z_interr)
    rc = pb.fn("dalton_solve_mv")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                  _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(pb.obs_ind),
                                  _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight), _host.ptr(pb.obs_var),
                                  _host.ptr(mean), _host.ptr(var), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "dalton.solve_mv")
    return pb.unbatch(mean), pb.unbatch(var)


def solve_sim(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
              obs_data=None, obs_times=None, obs_weight=None, obs_var=None, kalman_type="standard",
              prior_weight=None, prior_var=None, _z_interr=None, _z_smooth=None, _particle_offset=0, **params):
    r"""DALTON data-adaptive posterior draw (reference src/rodeo/inference/dalton.py:463-545)."""
    if key is None and _z_smooth is None:
        raise TypeError("solve_sim needs a PRNG key")
    pb = _adaptive(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, obs_data,
                   obs_times, obs_weight, obs_var, kalman_type, prior_weight, prior_var, params, _particle_offset)
    x = pb.empty(pb.B, pb.n_steps + 1, pb.nb, pb.p)
    n = getattr(pb.lib, f"rodeo_b200_dalton_solve_workspace_bytes_{pb.sfx}")(_lib.OP_SOLVE_SIM, ctypes.byref(pb.c))
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=_host.device())
    zi = None if _z_interr is None else pb.dev(_z_interr)
    zs = None if _z_smooth is None else pb.dev(_z_smooth)
    rc = pb.fn("dalton_solve_sim")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                   _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(zs),
                                   _host.ptr(pb.obs_ind), _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight),
                                   _host.ptr(pb.obs_var), _host.ptr(x), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "dalton.solve_sim")
    return pb.unbatch(x)
