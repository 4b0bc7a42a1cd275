"""Drop-in ``solve_mv`` / ``solve_sim`` (reference src/rodeo/solve.py:125-302) on the B200 kernels."""
import ctypes

import torch

from . import _host, _lib


def solve_mv(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
             kalman_type="standard", prior_weight=None, prior_var=None, _z_interr=None, **params):
    r"""Mean and variance of the stochastic ODE solver (reference src/rodeo/solve.py:208-302).

    Same arguments as the reference; ``theta`` / ``ode_init`` may carry a leading batch axis ``B``.

    Returns:
        mean (``[B,] n_steps+1, n_block, n_bstate``), var (``[B,] n_steps+1, n_block, n_bstate, n_bstate``)
        as CUDA tensors (float64, or float32 when ``theta`` / ``ode_init`` are float32).
    """
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params)
    N, dev = pb.n_steps, _host.device()
    mean = pb.empty(pb.B, N + 1, pb.nb, pb.p)
    var = pb.empty(pb.B, N + 1, pb.nb, pb.p, pb.p)
    zi = None if _z_interr is None else pb.dev(_z_interr)
    if kalman_type == "square-root":
        # prior_pars = (Q, cholesky(R)); `var` receives lower-triangular factors, as in the reference
        # (src/rodeo/kalmantv/square_root.py; docs/examples/higher_order.md:108-127)
        if pb.r_scale is not None or pb.prior_batch is not None:
            raise NotImplementedError('kalman_type="square-root" is compiled for a shared prior only')
        n = pb.lib.rodeo_b200_solve_mv_sqrt_workspace_bytes(ctypes.byref(pb.c))
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        rc = pb.fn("solve_mv_sqrt")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q),
                                                 _host.ptr(pb.R), _host.ptr(pb.x0), _host.ptr(pb.theta),
                                                 _host.ptr(zi), _host.ptr(mean), _host.ptr(var), _host.ptr(ws), n,
                                                 pb.stream())
        _lib.check(rc, "solve_mv[square-root]")
        return pb.unbatch(mean), pb.unbatch(var)
    ws, n = pb.workspace(_lib.OP_SOLVE_MV)
    rc = pb.fn("solve_mv")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                        _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(mean),
                                        _host.ptr(var), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "solve_mv")
    return pb.unbatch(mean), pb.unbatch(var)


def solve_sim(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
              kalman_type="standard", prior_weight=None, prior_var=None, _z_interr=None, _z_smooth=None,
              _particle_offset=0, **params):
    r"""One draw of the solution posterior per theta (reference src/rodeo/solve.py:125-205).

    Random draws come from a counter-based Philox generator keyed by ``(key, particle index, step)``; they are
    equal in distribution to the reference's, not bit-identical to JAX's threefry streams.
    ``_z_smooth`` / ``_z_interr`` inject standard normals instead (testing hook).

    Returns:
        ``[B,] n_steps+1, n_block, n_bstate`` float64 CUDA tensor.
    """
    if key is None and _z_smooth is None:
        raise TypeError("solve_sim needs a PRNG key")
    if kalman_type == "square-root":
        # the reference hands the square-root factor to multivariate_normal as if it were the covariance
        # (src/rodeo/solve.py:179 with kalman_funs = square_root): its draws have covariance (L L^T)^(1/2)
        raise NotImplementedError('solve_sim with kalman_type="square-root" is not built')
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params, particle_offset=_particle_offset)
    N, dev = pb.n_steps, _host.device()
    x = pb.empty(pb.B, N + 1, pb.nb, pb.p)
    ws, n = pb.workspace(_lib.OP_SOLVE_SIM)
    zi = None if _z_interr is None else pb.dev(_z_interr)
    zs = None if _z_smooth is None else pb.dev(_z_smooth)
    rc = pb.fn("solve_sim")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                         _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(zs),
                                         _host.ptr(x), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "solve_sim")
    return pb.unbatch(x)


def solve_sim_loglik(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars=None,
                     obs_data=None, obs_times=None, noise_sd=None, kalman_type="standard", prior_weight=None,
                     prior_var=None, return_draws=False, _z_interr=None, _z_smooth=None, _particle_offset=0, **params):
    r"""``solve_sim`` fused with the Gaussian observation log-likelihood of the draw: per theta,
    :math:`\sum_{i,k} \log N(y_{ik};\ X_{\mathrm{ind}(i), k, 0},\ \mathrm{sd}^2)` with ``ind = searchsorted(sim_times,
    obs_times)`` -- the ``logdensity_fn`` of the reference's pseudo-marginal MCMC walkthrough
    (docs/examples/parameter.md:333-354: ``rodeo.solve_sim``, ``Xt[obs_ind]``, ``fitz_loglik``) in ONE kernel.  The
    observation terms are accumulated inside the backward sweep; unless ``return_draws`` the trajectories are never
    written to memory.

    obs_data: ``(n_obs, n_block)``; obs_times: ``(n_obs,)`` sorted.  Returns ``loglik (B,)`` or ``(loglik, Xt)``.
    """
    if key is None and _z_smooth is None:
        raise TypeError("solve_sim_loglik needs a PRNG key")
    pb = _host.Problem(key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                       prior_weight, prior_var, kalman_type, params, particle_offset=_particle_offset)
    if pb.sfx != "f64" or kalman_type != "standard":
        raise NotImplementedError('solve_sim_loglik is compiled for float64, kalman_type="standard"')
    pb.set_obs(obs_data, obs_times)
    if pb.obs_data.dim() == 3 and pb.obs_data.shape[-1] == 1:
        pb.obs_data = pb.obs_data[..., 0].contiguous()
    if tuple(pb.obs_data.shape) != (pb.c.n_obs, pb.nb):
        raise ValueError("obs_data must have shape (n_obs, n_block)")
    if (pb.obs_ind_host[1:] < pb.obs_ind_host[:-1]).any():
        raise ValueError("obs_times must be sorted")
    N = pb.n_steps
    x = pb.empty(pb.B, N + 1, pb.nb, pb.p) if return_draws else None
    ll = pb.empty(pb.B)
    ws, n = pb.workspace(_lib.OP_SOLVE_SIM)
    zi = None if _z_interr is None else pb.dev(_z_interr)
    zs = None if _z_smooth is None else pb.dev(_z_smooth)
    rc = pb.lib.rodeo_b200_solve_sim_loglik_f64(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                                _host.ptr(pb.x0), _host.ptr(pb.theta), _host.ptr(zi), _host.ptr(zs),
                                                _host.ptr(pb.obs_ind), _host.ptr(pb.obs_data), float(noise_sd),
                                                _host.ptr(ll), _host.ptr(x), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "solve_sim_loglik")
    return (pb.unbatch(ll), pb.unbatch(x)) if return_draws else pb.unbatch(ll)
