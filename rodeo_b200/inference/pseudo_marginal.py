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


def init(position, logdensity_fn, rng_key):
    """reference src/rodeo/inference/pseudo_marginal.py:135-149"""
    position = _host.to_dev(position).clone()
    logdensity, auxdata = logdensity_fn(position, rng_key)
    return RWAState(position, logdensity.contiguous(), auxdata)


def _split3(rng_key):
    """three independent sub-keys (proposal, accept, log-density) of the step's key -- the role of
    jax.random.split(rng_key, 3) at reference :470 (not bit-compatible with threefry)"""
    k0, k1 = _host.parse_key(rng_key)
    mix = lambda a, b: ((k0 ^ a) & 0xFFFFFFFF, (k1 * 0x9E3779B1 + b) & 0xFFFFFFFF)
    return mix(0x243F6A88, 1), mix(0x85A308D3, 2), mix(0x13198A2E, 3)


def normal_random_walk(logdensity_fn, sigma):
    """RW-MH with a Gaussian proposal N(position, diag(sigma^2)) -- reference :175-189, :332-379, :452-483.

    ``logdensity_fn(position (C, d) CUDA tensor, key) -> (logdensity (C,), auxdata)``: typically
    ``rodeo_b200.solve_sim_loglik`` plus a log-prior, i.e. ONE fused kernel and no trajectories (``auxdata = None``).
    If it does return a tensor with a leading chain axis (e.g. the draws), only the ACCEPTED chains' rows are copied
    into the carried tensor (an indexed row copy; rejected rows are not touched).

    ``step(rng_key, state) -> (new_state, RWAInfo)``; ``rng_key`` is a uint32[2] key or an int (one per iteration,
    like the reference's ``jax.random.split`` keys).  Proposal and accept / reject are two small kernels
    (``rodeo_b200_rwmh_propose_f64`` / ``_accept_f64``); nothing returns to the host inside ``step``.  The state's
    ``position`` / ``logdensity`` (and ``auxdata``) buffers are updated IN PLACE and handed on to the new state.
    ``_z`` (C, d) / ``_u`` (C,) inject the proposal normals and the acceptance uniforms (testing hook).
    """
    sig = None
    lib = _lib.load()

    def init_fn(position, rng_key=None):
        return init(position, logdensity_fn, rng_key)

    def step_fn(rng_key, state, _z=None, _u=None):
        nonlocal sig
        pos, logd, aux = state
        C, d = pos.shape
        if sig is None:
            sig = _host.to_dev(torch.as_tensor(sigma, dtype=torch.float64).expand(d).contiguous())
        kp, ka, kl = _split3(rng_key)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        prop = torch.empty_like(pos)
        z = None if _z is None else _host.to_dev(_z)
        kbuf = (ctypes.c_uint32 * 2)(*kp)
        _lib.check(lib.rodeo_b200_rwmh_propose_f64(C, d, _host.ptr(pos), _host.ptr(sig), _host.ptr(z), kbuf, 0,
                                                   _host.ptr(prop), stream), "rwmh_propose")
        new_logd, new_aux = logdensity_fn(prop, kl)
        new_logd = new_logd.contiguous()
        acc = torch.empty(C, dtype=torch.int32, device=pos.device)
        p_accept = torch.empty(C, dtype=torch.float64, device=pos.device)
        u = None if _u is None else _host.to_dev(_u)
        kbuf = (ctypes.c_uint32 * 2)(*ka)
        _lib.check(lib.rodeo_b200_rwmh_accept_f64(C, d, _host.ptr(pos), _host.ptr(logd), _host.ptr(prop),
                                                  _host.ptr(new_logd), _host.ptr(u), kbuf, 0, _host.ptr(acc),
                                                  _host.ptr(p_accept), stream), "rwmh_accept")
        is_acc = acc.bool()
        if aux is not None and new_aux is not None:
            rows = is_acc.nonzero(as_tuple=True)[0]            # accepted chains only
            aux.index_copy_(0, rows, new_aux.index_select(0, rows))
        elif new_aux is not None:
            aux = new_aux
        new_state = RWAState(pos, logd, aux)
        return new_state, RWAInfo(p_accept, is_acc, new_state)     # reference :350: RWAInfo(p_accept, do_accept, new_state)

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
