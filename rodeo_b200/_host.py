"""Argument marshalling shared by the drop-in entry points (host side only; no arithmetic of the hot path).

Everything the reference takes un-batched and vmaps over (theta, ode_init) may carry a leading batch axis here;
an un-batched call is B = 1 and the outputs lose the batch axis again, so that a user of the reference sees the
reference's own shapes.  Tensors live on the current CUDA device; the kernels run on torch's current stream.
"""
import ctypes

import numpy as np
import torch

from . import _lib, interrogate as _interrogate, models as _models

_DT = torch.float64


def device():
    if not torch.cuda.is_available():
        raise _lib.RodeoError("rodeo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(x, dtype=_DT, dev=None):
    """array-like / numpy / torch (any device) -> contiguous tensor on the current CUDA device (or on `dev`)"""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    return t.to(device=device() if dev is None else dev, dtype=dtype, non_blocking=True).contiguous()


def on_host(*xs):
    """True when none of the given arrays lives on a CUDA device (NumPy arrays, lists, scalars, CPU tensors, None):
    such a call can go through the library's host-buffer entry points, which overlap the copies with the kernels."""
    return not any(isinstance(x, torch.Tensor) and x.is_cuda for x in xs)


def to_host(x, dtype=np.float64):
    """array-like / torch -> C-contiguous numpy (for the tiny W, Q, R that go in the parameter bank)"""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


def real_dtype(*xs):
    """float32 iff the ODE parameters / initial values are given in float32 (the reference's width follows
    jax_enable_x64, i.e. the dtype of the user's arrays); float64 otherwise."""
    for x in xs:
        dt = getattr(x, "dtype", None)
        if dt in (torch.float32, np.dtype(np.float32), np.float32):
            return torch.float32
    return torch.float64


def ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(t.data_ptr())


def parse_key(key):
    """jax-style PRNG key (uint32[2]), a python int, or None -> (k0, k1)"""
    if key is None:
        return 0, 0
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    k = np.asarray(key)
    if k.ndim == 0:
        v = int(k) & 0xFFFFFFFFFFFFFFFF
        return (v >> 32) & 0xFFFFFFFF, v & 0xFFFFFFFF
    k = k.astype(np.uint64).ravel()
    if k.size != 2:
        raise ValueError("key must be None, an int, or a uint32[2] PRNG key")
    return int(k[0]) & 0xFFFFFFFF, int(k[1]) & 0xFFFFFFFF


def prior_from(prior_pars, prior_weight, prior_var):
    """Accept both spellings: prior_pars=(Q, R) (reference HEAD, solve.py:208-212) and the <=1.1.2 keywords
    prior_weight=, prior_var= that BASELINE's north_star and the reference's examples/timings.py:30-35 use."""
    if prior_pars is not None:
        if prior_weight is not None or prior_var is not None:
            raise TypeError("give either prior_pars or prior_weight/prior_var, not both")
        prior_weight, prior_var = prior_pars
    if prior_weight is None or prior_var is None:
        raise TypeError("missing prior: pass prior_pars=(prior_weight, prior_var)")
    return prior_weight, prior_var


def kalman_id(kalman_type):
    # reference src/rodeo/solve.py:236-241
    if kalman_type == "standard":
        return _lib.KALMAN_STANDARD
    if kalman_type == "square-root":
        return _lib.KALMAN_SQUARE_ROOT
    raise NotImplementedError


class Problem:
    """Validated, device-resident inputs of one batched call."""

    def __init__(self, key, ode_fun, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars,
                 prior_weight, prior_var, kalman_type, params, particle_offset=0, host_inputs=False):
        # host_inputs: keep theta / ode_init / observation arrays as (zero-copy) CPU tensors for the *_host entry points
        self.where = torch.device("cpu") if host_inputs else device()
        self.model = _models.resolve(ode_fun)
        self.dtype = real_dtype(params.get("theta"), ode_init)
        self.sfx = "f32" if self.dtype == torch.float32 else "f64"
        self.esize = 4 if self.dtype == torch.float32 else 8
        npdt = np.float32 if self.dtype == torch.float32 else np.float64
        self.W = to_host(ode_weight, npdt)
        if self.W.ndim != 3:
            raise ValueError("ode_weight must have shape (n_block, n_bmeas, n_bstate)")
        self.nb, self.m, self.p = self.W.shape
        Qh, Rh = prior_from(prior_pars, prior_weight, prior_var)
        Qh, Rh = to_host(Qh), to_host(Rh)
        self.r_scale = None
        self.prior_batch = None
        if Rh.ndim == 4 or Qh.ndim == 4:
            # per-theta prior.  When it is a per-(theta, block) multiple of one shared matrix -- what ibm_init gives when
            # sigma is part of theta (R = sigma^2 R_1) -- the kernels keep the shared matrices in the constant bank and
            # take one scale per (theta, block).  Any other per-theta (Q, R) (the reference accepts whatever prior_pars
            # the user vmaps over, docs/examples/parameter.md:218-236) goes to the device as (B, n_block, p, p) arrays.
            Qb = Qh if Qh.ndim == 4 else np.broadcast_to(Qh, (Rh.shape[0],) + Qh.shape)
            Rb = Rh if Rh.ndim == 4 else np.broadcast_to(Rh, (Qh.shape[0],) + Rh.shape)
            if Rb.shape[0] == 0:                      # an empty batch: nothing to solve, any prior of the right shape
                Qb, Rb = np.zeros((1,) + Qb.shape[1:]), np.zeros((1,) + Rb.shape[1:])
            ref = Rb[0]
            with np.errstate(all="ignore"):
                scale = Rb[:, :, -1, -1] / ref[None, :, -1, -1]
            if np.all(Qb == Qb[:1]) and np.all(np.isfinite(scale)) and \
                    np.allclose(Rb, scale[:, :, None, None] * ref[None], rtol=1e-13, atol=0.0):
                self.r_scale_host, Qh, Rh = scale, Qb[0], ref
            else:
                self.prior_batch = (np.ascontiguousarray(Qb), np.ascontiguousarray(Rb))
                Qh, Rh = Qb[0], Rb[0]
        # the prior is built in float64 on the host (ibm_init) and rounded once to the compute type
        self.Q, self.R = to_host(Qh, npdt), to_host(Rh, npdt)
        if self.Q.shape != (self.nb, self.p, self.p) or self.R.shape != (self.nb, self.p, self.p):
            raise ValueError(f"prior matrices must have shape {(self.nb, self.p, self.p)} or (B, "
                             f"{self.nb}, {self.p}, {self.p}) (got {self.Q.shape}, {self.R.shape})")
        extra = set(params) - {"theta"}
        if extra:
            raise TypeError(f"unsupported ODE parameters {sorted(extra)}: device models take `theta` only")
        theta = params.get("theta")
        if theta is None:
            if self.model.n_theta:
                raise TypeError(f"model {self.model.name} needs theta=({self.model.n_theta},)")
            theta = np.zeros((1, 0))
        theta = to_dev(theta, self.dtype, self.where)
        x0 = to_dev(ode_init, self.dtype, self.where)
        self.batched = theta.ndim == 2 or x0.ndim == 3
        if theta.ndim == 1:
            theta = theta[None]
        if x0.ndim == 2:
            x0 = x0[None]
        if theta.ndim != 2 or x0.ndim != 3 or x0.shape[1:] != (self.nb, self.p):
            raise ValueError("theta must be (n_theta,) or (B, n_theta); ode_init (n_block, n_bstate) or (B, ...)")
        B = max(theta.shape[0], x0.shape[0])
        if theta.shape[0] not in (1, B) or x0.shape[0] not in (1, B):
            raise ValueError("theta and ode_init disagree on the batch size")
        self.B = B
        self.theta = theta.expand(B, theta.shape[1]).contiguous()
        self.x0 = x0.expand(B, self.nb, self.p).contiguous()
        self.n_steps = int(n_steps)
        self.t_min, self.t_max = float(t_min), float(t_max)
        k0, k1 = parse_key(key)
        c = _lib.RodeoProblem()
        c.B, c.particle_offset, c.n_steps = B, int(particle_offset), self.n_steps
        c.n_block, c.n_bstate, c.n_bmeas, c.n_theta = self.nb, self.p, self.m, self.theta.shape[1]
        c.model_id = self.model.model_id
        c.interrogate = _interrogate.resolve(interrogate, kalman_type)
        c.kalman_type = kalman_id(kalman_type)
        c.n_obs, c.n_bobs = 0, 0
        c.key[0], c.key[1] = k0, k1
        c.t_min, c.t_max = self.t_min, self.t_max
        c.user_wcol = self.model.wcol
        if getattr(self, "r_scale_host", None) is not None:
            if self.r_scale_host.shape[0] not in (1, B):
                raise ValueError("per-theta prior_var disagrees with the batch size")
            self.r_scale = to_dev(np.broadcast_to(self.r_scale_host, (B, self.nb)).copy(), self.dtype, self.where)
            c.prior_var_scale = self.r_scale.data_ptr()
            self.batched = True
        if self.prior_batch is not None:
            Qb, Rb = self.prior_batch
            if Qb.shape[0] not in (1, B):
                raise ValueError("per-theta prior disagrees with the batch size")
            if self.sfx != "f64":
                raise NotImplementedError("a general per-theta prior is compiled for float64 only")
            # self.Q / self.R are what the entry points receive: DEVICE arrays (B, n_block, p, p) with prior_batched set
            self.Q = to_dev(np.broadcast_to(Qb, (B,) + Qb.shape[1:]).copy(), self.dtype)
            self.R = to_dev(np.broadcast_to(Rb, (B,) + Rb.shape[1:]).copy(), self.dtype)
            c.prior_batched = 1
            self.batched = True
        self.c = c
        self.lib = _lib.load()

    # --- helpers -------------------------------------------------------------------------------------------------
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fn(self, name):
        """C-ABI entry point of this problem's arithmetic type, e.g. fn("solve_mv") -> rodeo_b200_solve_mv_f64"""
        return getattr(self.lib, f"rodeo_b200_{name}_{self.sfx}")

    def dev(self, x):
        return to_dev(x, self.dtype, self.where)

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.dtype, device=self.where)

    def workspace(self, op):
        n = self.lib.rodeo_b200_workspace_bytes(op, ctypes.byref(self.c), self.esize)
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device=device())
        return ws, n

    def set_obs(self, obs_data, obs_times, obs_weight=None, obs_var=None):
        """searchsorted(linspace(t_min,t_max,N+1), obs_times), left insertion, on the host exactly as the
        reference computes it (dalton.py:86-87), then the observation arrays on the device."""
        sim_times = np.linspace(self.t_min, self.t_max, self.n_steps + 1)
        ot = obs_times.detach().cpu().numpy() if isinstance(obs_times, torch.Tensor) else np.asarray(obs_times)
        ind = np.searchsorted(sim_times, ot.astype(np.float64)).astype(np.int32)
        self.obs_ind_host = ind
        self.obs_ind = torch.from_numpy(ind).to(self.where)
        self.c.n_obs = len(ind)
        self.obs_data = self.dev(obs_data)
        if obs_weight is not None:
            self.obs_weight, self.obs_var = self.dev(obs_weight), self.dev(obs_var)
            n_obs, nb, n_bobs, p = self.obs_weight.shape
            if (n_obs, nb, p) != (len(ind), self.nb, self.p) or self.obs_var.shape != (n_obs, nb, n_bobs, n_bobs) \
                    or self.obs_data.shape != (n_obs, nb, n_bobs):
                raise ValueError("obs_data (n_obs,n_block,n_bobs), obs_weight (n_obs,n_block,n_bobs,n_bstate), "
                                 "obs_var (n_obs,n_block,n_bobs,n_bobs) expected")
            self.c.n_bobs = n_bobs

    def unbatch(self, t):
        return t if self.batched else t[0]
