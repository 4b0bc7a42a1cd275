"""
CPU oracle for the rodeo probabilistic-ODE filtering hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a float64 NumPy restatement of the reference algorithm (mlysy/rodeo v1.1.3).  It is the
*checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``rodeo_b200/`` imports it, and
the product path has no CPU fallback.

PARITY PINNED TO THE REFERENCE'S OWN SOURCE, not to real JAX.  The reference is pure JAX and JAX is not installed
in this image (no network), and the reference repo ships no golden vectors.  But its unmodified source files run
over ``oracle/jaxshim`` -- a NumPy stand-in for the jax entry points they use (same LAPACK routines as jaxlib-CPU) --
and ``tests/golden/make_reference_golden.py`` commits what they compute as ``tests/golden/reference_vectors.npz``:
solve_mv (all four interrogations, chkrebtii on logged normals), solve_sim (logged normals, SVD factor), dalton,
fenrir, basic, dalton.solve_mv, fenrir.solve_mv, the square-root family, the Kalman primitives, the log-pdf's 1e-8
cut-off, ibm_init and first_order_pad, on FitzHugh-Nagumo (incl. the README's N = 800 walkthrough), Lorenz63 and the
second-order ODE; per-theta priors (a theta-dependent sigma, and arbitrary per-theta (Q, R)), two observation rows per
block (n_bobs = 2, correlated noise) and two measurement rows per block (n_bmeas = 2).  ``tests/test_reference_golden.py`` holds this oracle to those vectors at 1e-10 .. 1e-12 on the CPU
and the CUDA path at 1e-10 on the GPU.  What stays unpinned: XLA's own operation order (expected ~1e-13) and JAX's
threefry random streams (draws are compared on injected normals only).

Independent checks that do not involve the reference's code at all (``tests/test_oracle_*.py``):
  * the reference's own known-answer procedure for the Kalman primitives -- brute-force conditioning of the
    dense joint Gaussian of a random 3-step state-space model (reference tests/test_standard.py:18-200,
    tests/utils.py:24-63,117-215, tests/gauss_markov.py:30-125), restated in NumPy;
  * scan == for-loop composition / index conventions (reference tests/test_rodeofor.py:93-121,
    tests/ode_block_solve_for.py:81-235);
  * FitzHugh-Nagumo against scipy.integrate.odeint (reference tests/test_fitz.py:16-29);
  * the analytic solution of the docs' second-order ODE (reference docs/examples/higher_order.md:149-156);
  * dalton == fenrir == exact dense-Gaussian log p(Y | Z=0) for a linear ODE with interrogate_kramer
    (no reference test touches rodeo.inference; this three-way identity is independent of all three).

Every function is batched over a leading theta axis ``B`` (the reference is un-batched and relies on the
user's jax.vmap); all linear algebra goes through LAPACK via NumPy exactly where the reference goes through
jaxlib's LAPACK custom calls: ``np.linalg.solve`` (getrf/getrs) for ``rodeo.utils.solve_var``,
``np.linalg.eigh`` (syevd) for the log-pdf, ``np.linalg.cholesky`` / ``np.linalg.svd`` for the draws.

Layouts (C-contiguous, leading B):
    ode_weight (nb, m, p)      ode_init (B, nb, p)      prior Q, R (nb, p, p) or (B, nb, p, p)
    theta (B, n_theta)         obs_data (n_obs, nb, n_bobs)   obs_weight (n_obs, nb, n_bobs, p)
    obs_var (n_obs, nb, n_bobs, n_bobs)
"""
import math

import numpy as np

__all__ = [
    "ibm_init", "first_order_pad", "predict", "update", "forecast", "smooth_mv", "smooth_sim",
    "smooth_cond", "multivariate_normal_logpdf", "solve_filter", "solve_mv", "solve_sim", "basic",
    "fenrir", "dalton", "interrogate_kramer", "interrogate_chkrebtii", "interrogate_schober",
    "interrogate_rodeo", "MODELS", "obs_index", "psd_factor",
]


# ----------------------------------------------------------------------------------------------------
# prior  (reference src/rodeo/prior/ibm.py:21-88)
# ----------------------------------------------------------------------------------------------------

def _factorial(x):
    """exp(gammaln(x+1)) -- reference src/rodeo/prior/ibm.py:21-34 (not exactly integral in floating point)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for v in it:
        a = float(v) + 1.0
        # gammaln has poles at non-positive integers: +inf, as jax.scipy.special.gammaln returns
        if a <= 0.0 and a == math.floor(a):
            out[it.multi_index] = math.inf
        else:
            out[it.multi_index] = math.exp(math.lgamma(a))
    return out


def ibm_state(dt, q, sigma):
    """Q, R of the q-times integrated Brownian motion -- reference src/rodeo/prior/ibm.py:37-62."""
    I, J = np.meshgrid(np.arange(q + 1), np.arange(q + 1), indexing="ij", sparse=True)
    mesh = J - I
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        Q = np.nan_to_num(np.float64(dt) ** mesh.astype(np.float64) / _factorial(mesh), nan=0.0)
    mesh = (2.0 * q + 1.0) - I - J
    num = np.float64(dt) ** mesh
    den = mesh * _factorial(q - I) * _factorial(q - J)
    R = sigma ** 2 * num / den
    return Q, R


def ibm_init(dt, n_deriv, sigma):
    """Stacked per-block IBM prior -- reference src/rodeo/prior/ibm.py:65-88."""
    sigma = np.asarray(sigma, dtype=np.float64)
    n_block = len(sigma)
    Q1, R1 = ibm_state(dt, n_deriv - 1, 1)
    Q = np.repeat(Q1[None], n_block, axis=0)
    R = np.stack([sigma[b] ** 2 * R1 for b in range(n_block)])
    return Q, R


# ----------------------------------------------------------------------------------------------------
# ODE models: f(X, t, theta) -> (B, nb, m) and the block-diagonal Jacobian (B, nb, m, p)
# (what jax.jacfwd(ode_fun)[b, :, b] yields in reference src/rodeo/interrogate.py:75-79)
# ----------------------------------------------------------------------------------------------------

class OracleModel:
    def __init__(self, name, n_block, n_bstate, n_theta, fun, jac):
        self.name, self.n_block, self.n_bstate, self.n_theta = name, n_block, n_bstate, n_theta
        self.fun, self.jac = fun, jac


def _fn_fun(X, t, th):
    # reference README.md:92-99 / docs/examples/parameter.md (fitz_fun)
    a, b, c = th[:, 0], th[:, 1], th[:, 2]
    V, R = X[:, 0, 0], X[:, 1, 0]
    out = np.empty((X.shape[0], 2, 1))
    out[:, 0, 0] = c * (V - V * V * V / 3 + R)
    out[:, 1, 0] = -1 / c * (V - a + b * R)
    return out


def _fn_jac(X, t, th):
    a, b, c = th[:, 0], th[:, 1], th[:, 2]
    V = X[:, 0, 0]
    J = np.zeros((X.shape[0], 2, 1, X.shape[2]))
    J[:, 0, 0, 0] = c * (1 - V * V)
    J[:, 1, 0, 0] = -1 / c * b
    return J


def _lorenz_fun(X, t, th):
    # reference docs/examples/lorenz.md:95-101 ; theta = (rho, sigma, beta)
    rho, sig, beta = th[:, 0], th[:, 1], th[:, 2]
    x, y, z = X[:, 0, 0], X[:, 1, 0], X[:, 2, 0]
    out = np.empty((X.shape[0], 3, 1))
    out[:, 0, 0] = -sig * x + sig * y
    out[:, 1, 0] = rho * x - y - x * z
    out[:, 2, 0] = -beta * z + x * y
    return out


def _lorenz_jac(X, t, th):
    rho, sig, beta = th[:, 0], th[:, 1], th[:, 2]
    J = np.zeros((X.shape[0], 3, 1, X.shape[2]))
    J[:, 0, 0, 0] = -sig
    J[:, 1, 0, 0] = -1.0
    J[:, 2, 0, 0] = -beta
    return J


def _so_fun(X, t, th):
    # reference docs/examples/higher_order.md:47-58 generalised to theta = (omega, k): x'' = sin(omega t) - k x
    om, k = th[:, 0], th[:, 1]
    out = np.empty((X.shape[0], 1, 1))
    out[:, 0, 0] = np.sin(om * t) - k * X[:, 0, 0]
    return out


def _so_jac(X, t, th):
    J = np.zeros((X.shape[0], 1, 1, X.shape[2]))
    J[:, 0, 0, 0] = -th[:, 1]
    return J


def _hes1_fun(X, t, th):
    # reference examples/timings.py:253-262 (log-scale Hes1) ; theta = (a,b,c,d,e,f,g)
    P, M, H = np.exp(X[:, 0, 0]), np.exp(X[:, 1, 0]), np.exp(X[:, 2, 0])
    a, b, c, d, e, f, g = (th[:, i] for i in range(7))
    out = np.empty((X.shape[0], 3, 1))
    out[:, 0, 0] = -a * H + b * M / P - c
    out[:, 1, 0] = -d + e / (1 + P * P) / M
    out[:, 2, 0] = -a * P + f / (1 + P * P) / H - g
    return out


def _hes1_jac(X, t, th):
    P, M, H = np.exp(X[:, 0, 0]), np.exp(X[:, 1, 0]), np.exp(X[:, 2, 0])
    a, b, c, d, e, f, g = (th[:, i] for i in range(7))
    J = np.zeros((X.shape[0], 3, 1, X.shape[2]))
    J[:, 0, 0, 0] = -b * M / P                      # d/dlogP of b*M/P
    J[:, 1, 0, 0] = -e / (1 + P * P) / M            # d/dlogM of e/(1+P^2)/M
    J[:, 2, 0, 0] = -f / (1 + P * P) / H            # d/dlogH
    return J


def _seirah_fun(X, t, th):
    # reference examples/timings.py:339-351 ; theta = (b, r, alpha, D_e, D_I, D_q), N = S+E+I+R+A+H
    S, E, I, R, A, H = (X[:, i, 0] for i in range(6))
    b, r, alpha, D_e, D_I, D_q = (th[:, i] for i in range(6))
    N = S + E + I + R + A + H
    D_h = 30.0
    out = np.empty((X.shape[0], 6, 1))
    out[:, 0, 0] = -b * S * (I + alpha * A) / N
    out[:, 1, 0] = b * S * (I + alpha * A) / N - E / D_e
    out[:, 2, 0] = r * E / D_e - I / D_q - I / D_I
    out[:, 3, 0] = (I + A) / D_I + H / D_h
    out[:, 4, 0] = (1 - r) * E / D_e - A / D_I
    out[:, 5, 0] = I / D_q - H / D_h
    return out


def _seirah_jac(X, t, th):
    S, E, I, R, A, H = (X[:, i, 0] for i in range(6))
    b, r, alpha, D_e, D_I, D_q = (th[:, i] for i in range(6))
    N = S + E + I + R + A + H
    D_h = 30.0
    g = b * (I + alpha * A)
    J = np.zeros((X.shape[0], 6, 1, X.shape[2]))
    J[:, 0, 0, 0] = -g / N + g * S / (N * N)
    J[:, 1, 0, 0] = -b * S * (I + alpha * A) / (N * N) - 1 / D_e
    J[:, 2, 0, 0] = -1 / D_q - 1 / D_I
    J[:, 3, 0, 0] = 0.0
    J[:, 4, 0, 0] = -1 / D_I
    J[:, 5, 0, 0] = -1 / D_h
    return J


def _pair1b_fun(X, t, th):
    # tests/golden/make_reference_golden.py (pair_one_block): one block, two measured variables (n_bmeas = 2)
    out = np.empty((X.shape[0], 1, 2))
    out[:, 0, 0] = -th[:, 0] * X[:, 0, 0] + np.sin(t)
    out[:, 0, 1] = -th[:, 1] * X[:, 0, 3] * X[:, 0, 3]
    return out


def _pair1b_jac(X, t, th):
    J = np.zeros((X.shape[0], 1, 2, X.shape[2]))
    J[:, 0, 0, 0] = -th[:, 0]
    J[:, 0, 1, 3] = -2 * th[:, 1] * X[:, 0, 3]
    return J


MODELS = {
    "fitzhugh_nagumo": OracleModel("fitzhugh_nagumo", 2, 3, 3, _fn_fun, _fn_jac),
    "lorenz63": OracleModel("lorenz63", 3, 3, 3, _lorenz_fun, _lorenz_jac),
    "second_order_sin": OracleModel("second_order_sin", 1, 4, 2, _so_fun, _so_jac),
    "hes1": OracleModel("hes1", 3, 3, 7, _hes1_fun, _hes1_jac),
    "seirah": OracleModel("seirah", 6, 3, 6, _seirah_fun, _seirah_jac),
    "pair_one_block": OracleModel("pair_one_block", 1, 6, 3, _pair1b_fun, _pair1b_jac),
}


def first_order_pad(model, n_vars, n_deriv):
    """W and the initial-value helper -- reference src/rodeo/utils.py:80-102.

    ``ode_init(x0 (B, n_vars), t, theta (B, n_theta)) -> (B, n_vars, n_deriv)``.
    """
    def ode_init(x0, t, theta):
        x0 = np.asarray(x0, dtype=np.float64)
        X = np.zeros((x0.shape[0], n_vars, n_deriv))
        X[:, :, 0] = x0
        # the reference evaluates ode_fun(x0[:, None], t): only column 0 of X is visible to it
        X[:, :, 1] = model.fun(X[:, :, :1], t, np.asarray(theta, dtype=np.float64))[:, :, 0]
        return X

    W = np.zeros((n_vars, 1, n_deriv))
    W[:, :, 1] = 1.0
    return W, ode_init


# ----------------------------------------------------------------------------------------------------
# Kalman primitives, covariance form  (reference src/rodeo/kalmantv/standard.py)
# all arguments carry arbitrary leading batch axes
# ----------------------------------------------------------------------------------------------------

def _T(a):
    return np.swapaxes(a, -1, -2)


def _mv(A, x):
    return np.einsum("...ij,...j->...i", A, x)


def solve_var(V, B):
    """X = V^{-1} B by partial-pivot LU -- reference src/rodeo/utils.py:105-119 (jnp.linalg.solve)."""
    return np.linalg.solve(V, B)


def predict(mean_state_past, var_state_past, mean_state, wgt_state, var_state):
    """reference src/rodeo/kalmantv/standard.py:31-60"""
    mean_state_pred = _mv(wgt_state, mean_state_past) + mean_state
    var_state_pred = wgt_state @ var_state_past @ _T(wgt_state) + var_state
    return mean_state_pred, var_state_pred


def update(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas):
    """reference src/rodeo/kalmantv/standard.py:63-103"""
    mean_meas_pred = _mv(wgt_meas, mean_state_pred) + mean_meas
    var_meas_state_pred = wgt_meas @ var_state_pred
    var_meas_meas_pred = wgt_meas @ var_state_pred @ _T(wgt_meas) + var_meas
    var_state_meas_pred = var_state_pred @ _T(wgt_meas)
    var_state_temp = _T(solve_var(var_meas_meas_pred, _T(var_state_meas_pred)))
    mean_state_filt = mean_state_pred + _mv(var_state_temp, x_meas - mean_meas_pred)
    var_state_filt = var_state_pred - var_state_temp @ var_meas_state_pred
    return mean_state_filt, var_state_filt


def forecast(mean_state_pred, var_state_pred, mean_meas, wgt_meas, var_meas):
    """reference src/rodeo/kalmantv/standard.py:308-336"""
    mean_fore = _mv(wgt_meas, mean_state_pred) + mean_meas
    var_fore = wgt_meas @ var_state_pred @ _T(wgt_meas) + var_meas
    return mean_fore, var_fore


def _smooth(var_state_filt, var_state_pred, wgt_state):
    """reference src/rodeo/kalmantv/standard.py:160-177"""
    var_state_temp = var_state_filt @ _T(wgt_state)
    var_state_temp_tilde = _T(solve_var(var_state_pred, _T(var_state_temp)))
    return var_state_temp, var_state_temp_tilde


def smooth_mv(mean_state_next, var_state_next, mean_state_filt, var_state_filt,
              mean_state_pred, var_state_pred, wgt_state):
    """reference src/rodeo/kalmantv/standard.py:180-217"""
    _, G = _smooth(var_state_filt, var_state_pred, wgt_state)
    mean_state_smooth = mean_state_filt + _mv(G, mean_state_next - mean_state_pred)
    var_state_smooth = var_state_filt + G @ (var_state_next - var_state_pred) @ _T(G)
    return mean_state_smooth, var_state_smooth


def smooth_sim(x_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state):
    """reference src/rodeo/kalmantv/standard.py:220-255"""
    tmp, G = _smooth(var_state_filt, var_state_pred, wgt_state)
    mean_state_sim = mean_state_filt + _mv(G, x_state_next - mean_state_pred)
    var_state_sim = var_state_filt - G @ _T(tmp)
    return mean_state_sim, var_state_sim


def smooth_cond(mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state):
    """reference src/rodeo/kalmantv/standard.py:339-371"""
    tmp, A = _smooth(var_state_filt, var_state_pred, wgt_state)
    b = mean_state_filt - _mv(A, mean_state_pred)
    C = var_state_filt - A @ _T(tmp)
    return A, b, C


def multivariate_normal_logpdf(x, mean, cov):
    """Eigendecomposition log-pdf with the absolute 1e-8 eigenvalue cut-off.

    reference src/rodeo/utils.py:60-78: ``iw = ~isclose(w, 0, rtol=1e-300)`` keeps the default atol=1e-8,
    i.e. an eigenvalue is kept iff |w| > 1e-8 (+1e-300*0).
    """
    w, v = np.linalg.eigh(cov)
    z = np.einsum("...ji,...j->...i", v, x - mean)     # v.T @ (x - mean)
    z2 = z ** 2
    iw = ~np.isclose(w, 0.0, rtol=1e-300, atol=1e-8)
    w = np.where(iw, w, 1.0)
    val = z2 / w + np.log(w)
    return -0.5 * np.sum(np.where(iw, val, 0.0), axis=-1) - np.sum(iw, axis=-1) * 0.5 * np.log(2 * np.pi)


def _forecast_update(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas):
    """reference src/rodeo/inference/fenrir.py:40-81"""
    mean_fore, var_fore = forecast(mean_state_pred, var_state_pred, mean_meas, wgt_meas, var_meas)
    logdens = multivariate_normal_logpdf(x_meas, mean_fore, var_fore)
    mean_filt, var_filt = update(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas)
    return logdens, mean_filt, var_filt


# ----------------------------------------------------------------------------------------------------
# draws
# ----------------------------------------------------------------------------------------------------

def psd_factor(C, method):
    """A with A A^T = C.

    ``"svd"``      U sqrt(s): what jax.random.multivariate_normal(method='svd') uses (reference
                   src/rodeo/solve.py:179,182-186).
    ``"cholesky"`` np.linalg.cholesky: jax.random.multivariate_normal default (reference
                   src/rodeo/interrogate.py:30-34).
    ``"ldl"``      the device kernel's factor: un-pivoted L sqrt(D) with pivots d_j <= 0 clamped to a zero
                   column -- defined for singular PSD matrices; used for injected-normal parity tests.
    """
    if method == "svd":
        u, s, _ = np.linalg.svd(C)
        return u * np.sqrt(s[..., None, :])
    if method == "cholesky":
        return np.linalg.cholesky(C)
    if method == "ldl":
        C = np.array(C, dtype=np.float64, copy=True)
        p = C.shape[-1]
        L = np.zeros_like(C)
        for j in range(p):
            d = C[..., j, j].copy()
            for k in range(j):
                d = d - L[..., j, k] * L[..., j, k]
            pos = d > 0
            dj = np.sqrt(np.where(pos, d, 1.0))
            L[..., j, j] = np.where(pos, dj, 0.0)
            for i in range(j + 1, p):
                s = C[..., i, j].copy()
                for k in range(j):
                    s = s - L[..., i, k] * L[..., j, k]
                L[..., i, j] = np.where(pos, s / dj, 0.0)
        return L
    raise ValueError(method)


# ----------------------------------------------------------------------------------------------------
# interrogations  (reference src/rodeo/interrogate.py)
#   signature: (z, model, ode_weight, t, mean_state_pred, var_state_pred, theta) -> (wgt, mean, var)
#   `z` replaces the PRNG key: standard normals of shape (B, nb, p) (only chkrebtii consumes them)
# ----------------------------------------------------------------------------------------------------

def interrogate_kramer(z, model, ode_weight, t, mean_state_pred, var_state_pred, theta):
    """reference src/rodeo/interrogate.py:65-84 (block-diagonal Jacobian only)"""
    B, nb, p = mean_state_pred.shape
    m = ode_weight.shape[1]
    fun_meas = -model.fun(mean_state_pred, t, theta)
    jac = model.jac(mean_state_pred, t, theta)
    wgt_meas = -jac
    mean_meas = fun_meas + np.einsum("bnmp,bnp->bnm", jac, mean_state_pred)
    var_meas = np.zeros((B, nb, m, m))
    return wgt_meas, mean_meas, var_meas


def interrogate_schober(z, model, ode_weight, t, mean_state_pred, var_state_pred, theta):
    """reference src/rodeo/interrogate.py:50-62"""
    B, nb, p = mean_state_pred.shape
    m = ode_weight.shape[1]
    mean_meas = -model.fun(mean_state_pred, t, theta)
    return np.zeros((B,) + ode_weight.shape), mean_meas, np.zeros((B, nb, m, m))


def interrogate_rodeo(z, model, ode_weight, t, mean_state_pred, var_state_pred, theta):
    """reference src/rodeo/interrogate.py:87-115"""
    var_meas = ode_weight @ var_state_pred @ _T(ode_weight)
    mean_meas = -model.fun(mean_state_pred, t, theta)
    return np.zeros((mean_state_pred.shape[0],) + ode_weight.shape), mean_meas, var_meas


def interrogate_chkrebtii(z, model, ode_weight, t, mean_state_pred, var_state_pred, theta, factor="cholesky"):
    """reference src/rodeo/interrogate.py:13-47, kalman_type="standard" branch"""
    var_meas = ode_weight @ var_state_pred @ _T(ode_weight)
    A = psd_factor(var_state_pred, factor)
    x_state = mean_state_pred + _mv(A, z)
    mean_meas = -model.fun(x_state, t, theta)
    return np.zeros((mean_state_pred.shape[0],) + ode_weight.shape), mean_meas, var_meas


# ----------------------------------------------------------------------------------------------------
# solver  (reference src/rodeo/solve.py)
# ----------------------------------------------------------------------------------------------------

def _bq(Q, B):
    """broadcast a prior matrix to (B, nb, p, p)"""
    Q = np.asarray(Q, dtype=np.float64)
    return np.broadcast_to(Q, (B,) + Q.shape[-3:]) if Q.ndim == 3 else Q


def _step_time(t_min, t_max, n, n_steps):
    """reference src/rodeo/solve.py:74"""
    return t_min + (t_max - t_min) * (n + 1) / n_steps


def solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_weight, prior_var,
                 theta, z_interrogate=None, **ikw):
    """Forward pass -- reference src/rodeo/solve.py:31-122.

    Returns (mean_pred, var_pred, mean_filt, var_filt) with shapes (B, n_steps+1, nb, p[, p]).
    ``z_interrogate`` (B, n_steps, nb, p): the standard normals consumed by interrogate_chkrebtii.
    """
    ode_init = np.asarray(ode_init, dtype=np.float64)
    B, nb, p = ode_init.shape
    m = ode_weight.shape[1]
    Q, R = _bq(prior_weight, B), _bq(prior_var, B)
    x_meas = np.zeros((B, nb, m))
    mean_state = np.zeros((B, nb, p))
    mp = np.zeros((B, n_steps + 1, nb, p)); vp = np.zeros((B, n_steps + 1, nb, p, p))
    mf = np.zeros((B, n_steps + 1, nb, p)); vf = np.zeros((B, n_steps + 1, nb, p, p))
    mp[:, 0] = ode_init; mf[:, 0] = ode_init
    m_f, v_f = ode_init, np.zeros((B, nb, p, p))
    for n in range(n_steps):
        m_p, v_p = predict(m_f, v_f, mean_state, Q, R)
        z = None if z_interrogate is None else z_interrogate[:, n]
        wgt_meas, mean_meas, var_meas = interrogate(
            z, model, ode_weight, _step_time(t_min, t_max, n, n_steps), m_p, v_p, theta, **ikw)
        W_meas = ode_weight + wgt_meas
        m_f, v_f = update(m_p, v_p, x_meas, mean_meas, W_meas, var_meas)
        mp[:, n + 1], vp[:, n + 1], mf[:, n + 1], vf[:, n + 1] = m_p, v_p, m_f, v_f
    return mp, vp, mf, vf


def solve_mv(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
             z_interrogate=None, **ikw):
    """Posterior mean / variance -- reference src/rodeo/solve.py:208-302."""
    Qs, Rs = prior_pars
    mp, vp, mf, vf = solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate,
                                  Qs, Rs, theta, z_interrogate, **ikw)
    B = mf.shape[0]
    Q = _bq(Qs, B)
    ms = np.zeros_like(mf); vs = np.zeros_like(vf)
    ms[:, 0] = ode_init                       # row 0 is (ode_init, 0) verbatim, never smoothed
    ms[:, n_steps], vs[:, n_steps] = mf[:, n_steps], vf[:, n_steps]
    for t in range(n_steps - 1, 0, -1):
        ms[:, t], vs[:, t] = smooth_mv(ms[:, t + 1], vs[:, t + 1], mf[:, t], vf[:, t],
                                       mp[:, t + 1], vp[:, t + 1], Q)
    return ms, vs


def solve_sim(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
              z_smooth, z_interrogate=None, factor="svd", **ikw):
    """One posterior draw per theta -- reference src/rodeo/solve.py:125-205.

    ``z_smooth`` (B, n_steps+1, nb, p): standard normals; row ``n_steps`` feeds the terminal draw, rows
    ``1..n_steps-1`` the backward draws, row 0 is unused (x0 is known).
    """
    Qs, Rs = prior_pars
    mp, vp, mf, vf = solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate,
                                  Qs, Rs, theta, z_interrogate, **ikw)
    B = mf.shape[0]
    Q = _bq(Qs, B)
    xs = np.zeros_like(mf)
    xs[:, 0] = ode_init
    xs[:, n_steps] = mf[:, n_steps] + _mv(psd_factor(vf[:, n_steps], factor), z_smooth[:, n_steps])
    for t in range(n_steps - 1, 0, -1):
        m_sim, v_sim = smooth_sim(xs[:, t + 1], mf[:, t], vf[:, t], mp[:, t + 1], vp[:, t + 1], Q)
        xs[:, t] = m_sim + _mv(psd_factor(v_sim, factor), z_smooth[:, t])
    return xs


# ----------------------------------------------------------------------------------------------------
# likelihood layers  (reference src/rodeo/inference/{basic,fenrir,dalton}.py)
# ----------------------------------------------------------------------------------------------------

def obs_index(t_min, t_max, n_steps, obs_times):
    """searchsorted(linspace(t_min,t_max,N+1), obs_times), left insertion -- reference dalton.py:86-87"""
    sim_times = np.linspace(t_min, t_max, n_steps + 1)
    return np.searchsorted(sim_times, np.asarray(obs_times, dtype=np.float64)).astype(np.int64)


def basic(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
          obs_data, obs_times, obs_loglik, **kw):
    """reference src/rodeo/inference/basic.py:47-62 ; returns (loglik (B,), Xt)"""
    Xt, _ = solve_mv(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta, **kw)
    ind = obs_index(t_min, t_max, n_steps, obs_times)
    ode_data = Xt[:, ind]
    return obs_loglik(obs_data, ode_data, theta), Xt


def fenrir(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
           obs_data, obs_times, obs_weight, obs_var, z_interrogate=None, **ikw):
    """reference src/rodeo/inference/fenrir.py:86-328 ; returns loglik (B,)"""
    Qs, Rs = prior_pars
    mp, vp, mf, vf = solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate,
                                  Qs, Rs, theta, z_interrogate, **ikw)
    B, _, nb, p = mf.shape
    Q = _bq(Qs, B)
    obs_data = np.asarray(obs_data, dtype=np.float64)
    n_obs, _, n_bobs, _ = obs_weight.shape
    obs_ind = obs_index(t_min, t_max, n_steps, obs_times)
    obs_mean = np.zeros((B, nb, n_bobs))
    logdens = np.zeros(B)
    i = n_obs - 1
    bm, bv = mf[:, n_steps], vf[:, n_steps]
    # terminal point update  (fenrir.py:196-220)
    if obs_ind[i] >= n_steps:
        lp, bm, bv = _forecast_update(bm, bv, np.broadcast_to(obs_data[i], (B, nb, n_bobs)), obs_mean,
                                      obs_weight[i], obs_var[i])
        logdens += lp.sum(axis=1)
        i -= 1
    for t in range(n_steps - 1, -1, -1):          # fenrir.py:234, reverse scan over t = N-1 .. 0
        A, b, C = smooth_cond(mf[:, t], vf[:, t], mp[:, t + 1], vp[:, t + 1], Q)
        bm, bv = predict(bm, bv, b, A, C)
        # traced index semantics: negative i wraps NumPy-style
        if obs_ind[i] == t:
            lp, bm, bv = _forecast_update(bm, bv, np.broadcast_to(obs_data[i], (B, nb, n_bobs)), obs_mean,
                                          obs_weight[i], obs_var[i])
            logdens += lp.sum(axis=1)
            i -= 1
    return logdens


def dalton(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
           obs_data, obs_times, obs_weight, obs_var, z_interrogate=None, **ikw):
    """reference src/rodeo/inference/dalton.py:39-235 ; returns loglik (B,)

    ``z_interrogate`` (B, n_steps, 2, nb, p): normals for the joint ([:, :, 0]) and marginal ([:, :, 1])
    interrogations (dalton.py:226).
    """
    ode_init = np.asarray(ode_init, dtype=np.float64)
    B, nb, p = ode_init.shape
    m = ode_weight.shape[1]
    Qs, Rs = prior_pars
    Q, R = _bq(Qs, B), _bq(Rs, B)
    obs_data = np.asarray(obs_data, dtype=np.float64)
    n_obs, _, n_bobs, _ = obs_weight.shape
    obs_ind = obs_index(t_min, t_max, n_steps, obs_times)
    x_meas = np.zeros((B, nb, m)); obs_mean = np.zeros((B, nb, n_bobs)); mean_state = np.zeros((B, nb, p))

    logdens_zy = np.zeros(B); logdens_z = np.zeros(B)
    i = 0
    if obs_ind[0] == 0:                            # dalton.py:207-215
        for b in range(nb):
            logdens_zy += multivariate_normal_logpdf(
                obs_data[0, b], _mv(obs_weight[0, b], ode_init[:, b]) + obs_mean[:, b], obs_var[0, b])
        i = 1
    m_zy, v_zy = ode_init, np.zeros((B, nb, p, p))
    m_z, v_z = ode_init, np.zeros((B, nb, p, p))
    for n in range(n_steps):
        t = _step_time(t_min, t_max, n, n_steps)
        zj = None if z_interrogate is None else z_interrogate[:, n, 0]
        zm = None if z_interrogate is None else z_interrogate[:, n, 1]
        # joint filter (Z, Y)
        mp_zy, vp_zy = predict(m_zy, v_zy, mean_state, Q, R)
        wgt_meas, mean_meas, var_meas = interrogate(zj, model, ode_weight, t, mp_zy, vp_zy, theta, **ikw)
        W_meas = ode_weight + wgt_meas
        ic = min(i, n_obs - 1)                     # out-of-range traced gather clamps
        if n + 1 == obs_ind[ic]:
            D = np.broadcast_to(obs_weight[ic], (B, nb, n_bobs, p))
            wgt_obs = np.concatenate([W_meas, D], axis=2)
            mean_obs = np.concatenate([mean_meas, obs_mean], axis=2)
            var_obs = np.zeros((B, nb, m + n_bobs, m + n_bobs))
            var_obs[:, :, :m, :m] = var_meas
            var_obs[:, :, m:, m:] = obs_var[ic]
            x_obs = np.concatenate([x_meas, np.broadcast_to(obs_data[ic], (B, nb, n_bobs))], axis=2)
            lp, m_zy, v_zy = _forecast_update(mp_zy, vp_zy, x_obs, mean_obs, wgt_obs, var_obs)
            i += 1
        else:
            lp, m_zy, v_zy = _forecast_update(mp_zy, vp_zy, x_meas, mean_meas, W_meas, var_meas)
        logdens_zy += lp.sum(axis=1)
        # marginal filter (Z)
        mp_z, vp_z = predict(m_z, v_z, mean_state, Q, R)
        wgt_meas, mean_meas, var_meas = interrogate(zm, model, ode_weight, t, mp_z, vp_z, theta, **ikw)
        W_meas = ode_weight + wgt_meas
        lp, m_z, v_z = _forecast_update(mp_z, vp_z, x_meas, mean_meas, W_meas, var_meas)
        logdens_z += lp.sum(axis=1)
    return logdens_zy - logdens_z


# ----------------------------------------------------------------------------------------------------
# data-adaptive solvers  (reference src/rodeo/inference/dalton.py:242-545)
# ----------------------------------------------------------------------------------------------------

def dalton_solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_weight, prior_var,
                        theta, obs_data, obs_times, obs_weight, obs_var, z_interrogate=None, **ikw):
    """Forward pass that also conditions on the observations -- reference dalton.py:242-371."""
    ode_init = np.asarray(ode_init, dtype=np.float64)
    B, nb, p = ode_init.shape
    m = ode_weight.shape[1]
    Q, R = _bq(prior_weight, B), _bq(prior_var, B)
    obs_data = np.asarray(obs_data, dtype=np.float64)
    n_obs, _, n_bobs, _ = obs_weight.shape
    obs_ind = obs_index(t_min, t_max, n_steps, obs_times)
    x_meas = np.zeros((B, nb, m)); obs_mean = np.zeros((B, nb, n_bobs)); mean_state = np.zeros((B, nb, p))
    mp = np.zeros((B, n_steps + 1, nb, p)); vp = np.zeros((B, n_steps + 1, nb, p, p))
    mf = np.zeros((B, n_steps + 1, nb, p)); vf = np.zeros((B, n_steps + 1, nb, p, p))
    mp[:, 0] = ode_init; mf[:, 0] = ode_init
    m_f, v_f = ode_init, np.zeros((B, nb, p, p))
    i = 1 if obs_ind[0] == 0 else 0                      # dalton.py:352
    for n in range(n_steps):
        m_p, v_p = predict(m_f, v_f, mean_state, Q, R)
        z = None if z_interrogate is None else z_interrogate[:, n]
        wgt_meas, mean_meas, var_meas = interrogate(z, model, ode_weight, _step_time(t_min, t_max, n, n_steps),
                                                    m_p, v_p, theta, **ikw)
        W_meas = ode_weight + wgt_meas
        ic = min(i, n_obs - 1)
        if n + 1 == obs_ind[ic]:
            D = np.broadcast_to(obs_weight[ic], (B, nb, n_bobs, p))
            wgt_obs = np.concatenate([W_meas, D], axis=2)
            mean_obs = np.concatenate([mean_meas, obs_mean], axis=2)
            var_obs = np.zeros((B, nb, m + n_bobs, m + n_bobs))
            var_obs[:, :, :m, :m] = var_meas
            var_obs[:, :, m:, m:] = obs_var[ic]
            x_obs = np.concatenate([x_meas, np.broadcast_to(obs_data[ic], (B, nb, n_bobs))], axis=2)
            m_f, v_f = update(m_p, v_p, x_obs, mean_obs, wgt_obs, var_obs)
            i += 1
        else:
            m_f, v_f = update(m_p, v_p, x_meas, mean_meas, W_meas, var_meas)
        mp[:, n + 1], vp[:, n + 1], mf[:, n + 1], vf[:, n + 1] = m_p, v_p, m_f, v_f
    return mp, vp, mf, vf


def dalton_solve_mv(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
                    obs_data, obs_times, obs_weight, obs_var, **kw):
    """reference dalton.py:374-460 (same backward pass as solve_mv)"""
    Qs, Rs = prior_pars
    mp, vp, mf, vf = dalton_solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, Qs, Rs,
                                         theta, obs_data, obs_times, obs_weight, obs_var, **kw)
    Q = _bq(Qs, mf.shape[0])
    ms = np.zeros_like(mf); vs = np.zeros_like(vf)
    ms[:, 0] = ode_init
    ms[:, n_steps], vs[:, n_steps] = mf[:, n_steps], vf[:, n_steps]
    for t in range(n_steps - 1, 0, -1):
        ms[:, t], vs[:, t] = smooth_mv(ms[:, t + 1], vs[:, t + 1], mf[:, t], vf[:, t], mp[:, t + 1], vp[:, t + 1], Q)
    return ms, vs


def dalton_solve_sim(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
                     obs_data, obs_times, obs_weight, obs_var, z_smooth, factor="svd", **kw):
    """reference dalton.py:463-545 (same backward pass as solve_sim)"""
    Qs, Rs = prior_pars
    mp, vp, mf, vf = dalton_solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, Qs, Rs,
                                         theta, obs_data, obs_times, obs_weight, obs_var, **kw)
    Q = _bq(Qs, mf.shape[0])
    xs = np.zeros_like(mf)
    xs[:, 0] = ode_init
    xs[:, n_steps] = mf[:, n_steps] + _mv(psd_factor(vf[:, n_steps], factor), z_smooth[:, n_steps])
    for t in range(n_steps - 1, 0, -1):
        m_sim, v_sim = smooth_sim(xs[:, t + 1], mf[:, t], vf[:, t], mp[:, t + 1], vp[:, t + 1], Q)
        xs[:, t] = m_sim + _mv(psd_factor(v_sim, factor), z_smooth[:, t])
    return xs


def magi_logdens(ode_state, n_active, prior_pars):
    """reference src/rodeo/inference/magi.py:6-99 after its ``ode_expand`` call: ``ode_state`` (B, N+1, nb, p) is the
    expanded solution process.  Noise-free observation of the first ``n_active`` state entries of every block; the
    log-density is the Cholesky-type one of jax.scipy.stats.multivariate_normal.logpdf (no eigenvalue cut-off)."""
    X = np.asarray(ode_state, dtype=np.float64)
    B, N1, nb, p = X.shape
    Q, R = _bq(prior_pars[0], B), _bq(prior_pars[1], B)
    W = np.broadcast_to(np.eye(n_active, p), (B, nb, n_active, p))
    zero_m, zero_v = np.zeros((B, nb, n_active)), np.zeros((B, nb, n_active, n_active))
    mean_state = np.zeros((B, nb, p))
    m, v = X[:, 0], np.zeros((B, nb, p, p))
    total = np.zeros(B)
    for n in range(1, N1):
        x_meas = X[:, n, :, :n_active]
        mp, vp = predict(m, v, mean_state, Q, R)
        mf, vf = forecast(mp, vp, zero_m, W, zero_v)
        try:
            L = np.linalg.cholesky(0.5 * (vf + _T(vf)))      # jnp.linalg.cholesky symmetrises its input
        except np.linalg.LinAlgError:                        # ... and returns NaN where LAPACK reports failure
            return np.full(B, np.nan)
        z = np.linalg.solve(L, (x_meas - mf)[..., None])[..., 0]
        logp = -0.5 * np.sum(z * z, -1) - np.sum(np.log(np.diagonal(L, axis1=-2, axis2=-1)), -1) \
            - 0.5 * n_active * math.log(2 * math.pi)
        total += logp.sum(-1)
        m, v = update(mp, vp, x_meas, zero_m, W, zero_v)
    return total


# ----------------------------------------------------------------------------------------------------
# square-root Kalman family  (reference src/rodeo/kalmantv/square_root.py, src/rodeo/utils.py:10-24)
# variances are carried as lower-triangular factors L (var = L L^T); only L L^T is comparable across
# implementations (QR leaves the signs of R's diagonal free)
# ----------------------------------------------------------------------------------------------------

def add_sqrt(sqrt_A, sqrt_B):
    """R^T of qr(vstack([sqrt_A^T, sqrt_B^T])) -- reference src/rodeo/utils.py:10-24; batched over leading axes."""
    stacked = np.concatenate([_T(sqrt_A), _T(sqrt_B)], axis=-2)
    R = np.linalg.qr(stacked, mode="r")
    return _T(R)


def _solve_tri(L, B, lower, trans=False):
    import scipy.linalg
    L = np.asarray(L); B = np.asarray(B)
    lead = np.broadcast_shapes(L.shape[:-2], B.shape[:-2])
    Lb = np.broadcast_to(L, lead + L.shape[-2:]).reshape((-1,) + L.shape[-2:])
    Bb = np.broadcast_to(B, lead + B.shape[-2:]).reshape((-1,) + B.shape[-2:])
    out = np.stack([scipy.linalg.solve_triangular(l, b, lower=lower, trans="T" if trans else "N")
                    for l, b in zip(Lb, Bb)])
    return out.reshape(lead + B.shape[-2:])


def sqrt_predict(mean_state_past, var_state_past, mean_state, wgt_state, var_state):
    """reference square_root.py:30-59"""
    return _mv(wgt_state, mean_state_past) + mean_state, add_sqrt(wgt_state @ var_state_past, var_state)


def sqrt_update(mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas):
    """reference square_root.py:62-103"""
    mean_meas_pred = _mv(wgt_meas, mean_state_pred) + mean_meas
    var_meas_meas_pred = add_sqrt(wgt_meas @ var_state_pred, var_meas)
    inter = _solve_tri(var_meas_meas_pred, wgt_meas, lower=True)
    inter = inter @ var_state_pred @ _T(var_state_pred)
    var_state_temp = _T(_solve_tri(_T(var_meas_meas_pred), inter, lower=False))
    mean_state_filt = mean_state_pred + _mv(var_state_temp, x_meas - mean_meas_pred)
    var_state_filt = add_sqrt(var_state_pred - (var_state_temp @ wgt_meas) @ var_state_pred,
                              var_state_temp @ var_meas)
    return mean_state_filt, var_state_filt


def _sqrt_smooth(var_state_filt, var_state_pred, wgt_state):
    """reference square_root.py:160-178"""
    variance_state_filt = var_state_filt @ _T(var_state_filt)
    inter = _solve_tri(var_state_pred, wgt_state, lower=True) @ variance_state_filt
    return _T(_solve_tri(_T(var_state_pred), inter, lower=False))


def sqrt_smooth_mv(mean_state_next, var_state_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred,
                   wgt_state, var_state):
    """reference square_root.py:181-222"""
    G = _sqrt_smooth(var_state_filt, var_state_pred, wgt_state)
    mean_state_smooth = mean_state_filt + _mv(G, mean_state_next - mean_state_pred)
    J = np.eye(G.shape[-1]) - G @ wgt_state
    lead = np.broadcast_shapes(var_state_next.shape[:-2], np.asarray(var_state).shape[:-2])
    hs = np.concatenate([np.broadcast_to(var_state_next, lead + var_state_next.shape[-2:]),
                         np.broadcast_to(var_state, lead + np.asarray(var_state).shape[-2:])], axis=-1)
    return mean_state_smooth, add_sqrt(G @ hs, J @ var_state_filt)


def interrogate_chkrebtii_sqrt(z, model, ode_weight, t, mean_state_pred, var_state_pred, theta):
    """reference src/rodeo/interrogate.py:36-47, kalman_type="square-root" branch, restated literally:
    var_meas = W L (an (m, p) block, not (m, m)) and x_state = mean + var_meas @ z, whose (m,) = (1,) result
    BROADCASTS over all p state entries of the block."""
    var_meas = ode_weight @ var_state_pred                               # (B, nb, m, p)
    shift = np.einsum("bnmp,bnp->bnm", var_meas, z)                     # (B, nb, m)
    x_state = mean_state_pred + shift[..., :1]                          # m == 1: broadcast over p
    mean_meas = -model.fun(x_state, t, theta)
    return np.zeros((mean_state_pred.shape[0],) + ode_weight.shape), mean_meas, var_meas


def solve_mv_sqrt(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
                  z_interrogate=None):
    """solve_mv with kalman_type="square-root" -- reference src/rodeo/solve.py:208-302 with
    kalman_funs = square_root.  prior_pars = (Q, lower Cholesky factor of R).  Returns (mean, L)."""
    ode_init = np.asarray(ode_init, dtype=np.float64)
    B, nb, p = ode_init.shape
    m = ode_weight.shape[1]
    Q, Rh = _bq(prior_pars[0], B), _bq(prior_pars[1], B)
    x_meas = np.zeros((B, nb, m)); mean_state = np.zeros((B, nb, p))
    N = n_steps
    mp = np.zeros((B, N + 1, nb, p)); Lp = np.zeros((B, N + 1, nb, p, p))
    mf = np.zeros((B, N + 1, nb, p)); Lf = np.zeros((B, N + 1, nb, p, p))
    mp[:, 0] = ode_init; mf[:, 0] = ode_init
    m_f, l_f = ode_init, np.zeros((B, nb, p, p))
    for n in range(N):
        m_p, l_p = sqrt_predict(m_f, l_f, mean_state, Q, Rh)
        z = None if z_interrogate is None else z_interrogate[:, n]
        wgt_meas, mean_meas, var_meas = interrogate(z, model, ode_weight, _step_time(t_min, t_max, n, N), m_p, l_p,
                                                    theta)
        W_meas = ode_weight + wgt_meas
        m_f, l_f = sqrt_update(m_p, l_p, x_meas, mean_meas, W_meas, var_meas)
        mp[:, n + 1], Lp[:, n + 1], mf[:, n + 1], Lf[:, n + 1] = m_p, l_p, m_f, l_f
    ms = np.zeros_like(mf); Ls = np.zeros_like(Lf)
    ms[:, 0] = ode_init
    ms[:, N], Ls[:, N] = mf[:, N], Lf[:, N]
    for t in range(N - 1, 0, -1):
        ms[:, t], Ls[:, t] = sqrt_smooth_mv(ms[:, t + 1], Ls[:, t + 1], mf[:, t], Lf[:, t], mp[:, t + 1],
                                            Lp[:, t + 1], Q, Rh)
    return ms, Ls


# ----------------------------------------------------------------------------------------------------
# fenrir solver  (reference src/rodeo/inference/fenrir.py:86-259 _backward stacks, :333-457 _smooth_mv / solve_mv)
# ----------------------------------------------------------------------------------------------------

def fenrir_solve_mv(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, prior_pars, theta,
                    obs_data, obs_times, obs_weight, obs_var, z_interrogate=None, **ikw):
    """Posterior mean / variance of p(X_{0:N} | Z_{1:N}, Y_{0:M}) by the Fenrir construction: forward ODE filter,
    backward Markov chain filtered with the observations (fenrir.py:86-259), then an RTS pass over that chain forward
    in time (fenrir.py:333-401).  Returns (mean, var) of shapes (B, N+1, nb, p[, p])."""
    Qs, Rs = prior_pars
    mp, vp, mf, vf = solve_filter(model, ode_weight, ode_init, t_min, t_max, n_steps, interrogate, Qs, Rs, theta,
                                  z_interrogate, **ikw)
    B, _, nb, p = mf.shape
    N = n_steps
    Q = _bq(Qs, B)
    obs_data = np.asarray(obs_data, dtype=np.float64)
    n_obs, _, n_bobs, _ = obs_weight.shape
    obs_ind = obs_index(t_min, t_max, N, obs_times)
    obs_mean = np.zeros((B, nb, n_bobs))
    # stacks over t = 0..N (entry N = terminal point)
    bmp = np.zeros((B, N + 1, nb, p)); bvp = np.zeros((B, N + 1, nb, p, p))
    bmf = np.zeros((B, N + 1, nb, p)); bvf = np.zeros((B, N + 1, nb, p, p))
    A_all = np.zeros((B, N, nb, p, p))
    i = n_obs - 1
    bm, bv = mf[:, N], vf[:, N]
    bmp[:, N], bvp[:, N] = bm, bv                        # "state_pred" terminal entry = forward filt[N]  (fenrir.py:240-245)
    if obs_ind[i] >= N:
        _, bm, bv = _forecast_update(bm, bv, np.broadcast_to(obs_data[i], (B, nb, n_bobs)), obs_mean, obs_weight[i],
                                     obs_var[i])
        i -= 1
    bmf[:, N], bvf[:, N] = bm, bv
    for t in range(N - 1, -1, -1):
        A, b, C = smooth_cond(mf[:, t], vf[:, t], mp[:, t + 1], vp[:, t + 1], Q)
        bm, bv = predict(bm, bv, b, A, C)
        bmp[:, t], bvp[:, t] = bm, bv
        A_all[:, t] = A
        if obs_ind[i] == t:
            _, bm, bv = _forecast_update(bm, bv, np.broadcast_to(obs_data[i], (B, nb, n_bobs)), obs_mean, obs_weight[i],
                                         obs_var[i])
            i -= 1
        bmf[:, t], bvf[:, t] = bm, bv
    # forward-in-time smoothing of the backward chain  (fenrir.py:333-401)
    ms = np.zeros_like(bmf); vs = np.zeros_like(bvf)
    ms[:, 0:2], vs[:, 0:2] = bmf[:, 0:2], bvf[:, 0:2]
    for k in range(N - 1):                               # scan index; produces row k + 2
        ms[:, k + 2], vs[:, k + 2] = smooth_mv(ms[:, k + 1], vs[:, k + 1], bmf[:, k + 2], bvf[:, k + 2],
                                               bmp[:, k + 1], bvp[:, k + 1], A_all[:, k + 1])
    return ms, vs


# ---------------------------------------------------------------------------------------------------------------------
# pseudo-marginal random-walk Metropolis-Hastings, many chains  (reference src/rodeo/inference/pseudo_marginal.py)
# ---------------------------------------------------------------------------------------------------------------------
def gauss_obs_loglik(Xt, obs_ind, obs_data, noise_sd):
    """sum_{i,k} log N(obs_data[i,k]; Xt[:, obs_ind[i], k, 0], noise_sd^2): the `fitz_loglik` of the reference's
    parameter-inference walkthrough (docs/examples/parameter.md:192-205) applied to Xt[obs_ind]"""
    ind = np.clip(np.asarray(obs_ind), 0, Xt.shape[1] - 1)
    d = np.asarray(obs_data)[None] - Xt[..., 0][:, ind]
    return np.sum(-0.5 * np.log(2 * np.pi) - np.log(noise_sd) - 0.5 * d * d / noise_sd ** 2, axis=(1, 2))


def rwmh_step(position, logdensity, logdensity_fn, sigma, z, u):
    """One RW-MH step for every chain with injected proposal normals z (C, d) and acceptance uniforms u (C,):
    new_position = position + sigma z (pseudo_marginal.py:175-189, diagonal sigma); proposed log-density from
    `logdensity_fn(new_position)` (:473); log_p = new - old for the symmetric proposal with NaN -> -inf, p_accept =
    min(1, exp(log_p)), accept iff u < p_accept (blackjax compute_asymmetric_acceptance_ratio /
    static_binomial_sampling, as called at :476-479).  Returns (position, logdensity, accepted, p_accept)."""
    prop = position + np.asarray(sigma) * z
    new_ld = logdensity_fn(prop)
    lp = new_ld - logdensity
    lp = np.where(np.isnan(lp), -np.inf, lp)
    with np.errstate(over="ignore"):
        pa = np.minimum(np.exp(lp), 1.0)
    acc = u < pa
    return np.where(acc[:, None], prop, position), np.where(acc, new_ld, logdensity), acc, pa
