"""Device-resident pseudo-marginal random-walk Metropolis-Hastings for MANY chains at once.

Mirrors the API of the reference's ``rodeo.inference.pseudo_marginal`` (src/rodeo/inference/pseudo_marginal.py, a
blackjax-style RW-MH whose ``logdensity_fn(position, key) -> (logdensity, auxdata)`` carries auxiliary data, e.g.
the sampled ODE solution): ``init``, ``normal_random_walk(logdensity_fn, sigma) -> SamplingAlgorithm(init, step)``,
``RWAState``, ``RWAInfo``.  Differences: ``position`` is a CUDA tensor ``(n_chains, dim)`` -- every chain is one
theta of the batched kernels -- and proposal, log-density (a batched ``solve_sim`` + fused Gaussian observation
log-likelihood), accept/reject all stay on the device: nothing returns to the host inside ``step``
(the reference evaluates ONE solve_sim per iteration, docs/examples/parameter.md:333-396).

Only the heavy inner call runs in hand-written kernels; proposal and accept/reject are O(n_chains * dim) elementwise
torch ops on the current stream.
"""
import ctypes
from typing import Any, Callable, NamedTuple

import torch

from .. import _host, _lib


class RWAState(NamedTuple):
    """reference src/rodeo/inference/pseudo_marginal.py:103-116"""
    position: Any
    logdensity: Any
    auxdata: Any


class RWAInfo(NamedTuple):
    """reference src/rodeo/inference/pseudo_marginal.py:119-132"""
    acceptance_rate: Any
    is_accepted: Any
    proposal: Any


class SamplingAlgorithm(NamedTuple):
    init: Callable
    step: Callable


def _generator(key):
    k0, k1 = _host.parse_key(key)
    g = torch.Generator(device=_host.device())
    g.manual_seed(((k0 << 32) | k1) & 0x7FFFFFFFFFFFFFFF)
    return g


def init(position, logdensity_fn, rng_key):
    """reference src/rodeo/inference/pseudo_marginal.py:135-149"""
    position = _host.to_dev(position)
    logdensity, auxdata = logdensity_fn(position, rng_key)
    return RWAState(position, logdensity, auxdata)


def normal_random_walk(logdensity_fn, sigma):
    """RW-MH with a Gaussian proposal N(position, diag(sigma^2)) -- reference :175-189, :332-379, :452-483.

    ``logdensity_fn(position (C, d) CUDA tensor, key) -> (logdensity (C,), auxdata)``; ``auxdata`` may be a tensor with
    a leading chain axis (it is carried per chain through accept/reject) or ``None``.
    ``step(rng_key, state) -> (new_state, RWAInfo)``; ``rng_key`` is a uint32[2] key or an int (one per iteration,
    like the reference's ``jax.random.split`` keys).
    """
    sig = None

    def init_fn(position, rng_key=None):
        return init(position, logdensity_fn, rng_key)

    def step_fn(rng_key, state):
        nonlocal sig
        pos, logd, aux = state
        if sig is None:
            sig = _host.to_dev(sigma, pos.dtype)
        g = _generator(rng_key)
        # key_proposal / key_accept / key_logdensity of the reference (:470): three independent streams
        prop = pos + sig * torch.randn(pos.shape, dtype=pos.dtype, device=pos.device, generator=g)
        logu = torch.log(torch.rand(pos.shape[0], dtype=pos.dtype, device=pos.device, generator=g))
        k0, k1 = _host.parse_key(rng_key)
        new_logd, new_aux = logdensity_fn(prop, (k0 ^ 0x9E3779B9, k1 ^ 0x85EBCA6B))
        log_p = new_logd - logd                                  # symmetric proposal (:438-443)
        log_p = torch.where(torch.isnan(log_p), torch.full_like(log_p, -float("inf")), log_p)
        acc = logu < log_p
        new_pos = torch.where(acc[:, None], prop, pos)
        out_logd = torch.where(acc, new_logd, logd)
        if aux is None or new_aux is None:
            out_aux = new_aux
        else:
            out_aux = torch.where(acc.view((-1,) + (1,) * (new_aux.dim() - 1)), new_aux, aux)
        p_accept = torch.clamp(torch.exp(log_p), max=1.0)
        return RWAState(new_pos, out_logd, out_aux), RWAInfo(p_accept, acc, RWAState(prop, new_logd, new_aux))

    return SamplingAlgorithm(init_fn, step_fn)


def gauss_obs_loglik(Xt, obs_ind, obs_data, noise_sd):
    """``sum_{i,k} log N(obs_data[i,k]; Xt[:, obs_ind[i], k, 0], noise_sd^2)`` per theta, fused gather + reduction on
    the device (the ``fitz_loglik`` of docs/examples/parameter.md:192-205 applied to ``Xt[obs_ind]``).

    Xt: ``(B, n_steps+1, n_block, n_bstate)`` CUDA float64; obs_ind: int tensor/array ``(n_obs,)``;
    obs_data: ``(n_obs, n_block)``.  Returns ``(B,)``.
    """
    lib = _lib.load()
    Xt = Xt.contiguous()
    B, n1, nb, p = Xt.shape
    ind = torch.as_tensor(obs_ind, dtype=torch.int32).to(Xt.device).contiguous()
    y = _host.to_dev(obs_data).contiguous()
    c = _lib.RodeoProblem()
    c.B, c.n_steps, c.n_block, c.n_bstate, c.n_obs = B, n1 - 1, nb, p, ind.numel()
    out = torch.empty((B,), dtype=torch.float64, device=Xt.device)
    rc = lib.rodeo_b200_gauss_obs_loglik_f64(ctypes.byref(c), _host.ptr(Xt), _host.ptr(ind), _host.ptr(y),
                                             float(noise_sd), _host.ptr(out),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "gauss_obs_loglik")
    return out
