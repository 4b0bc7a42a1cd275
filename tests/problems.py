"""Seeded synthetic problem set-ups of BASELINE.json's configs, shared by tests, smoke and bench.

Pure NumPy: no product code, no oracle.  Every array follows the reference's layouts.
"""
import numpy as np


def _ibm(dt, n_deriv, sigma):
    # local copy of the closed form so this module depends on neither the product nor the oracle
    import math
    q = n_deriv - 1
    Q = np.zeros((n_deriv, n_deriv)); R = np.zeros((n_deriv, n_deriv))
    fact = lambda k: math.exp(math.lgamma(k + 1.0))
    for i in range(n_deriv):
        for j in range(n_deriv):
            if j >= i:
                Q[i, j] = dt ** (j - i) / fact(j - i)
            e = 2.0 * q + 1.0 - i - j
            R[i, j] = dt ** e / (e * fact(q - i) * fact(q - j))
    sigma = np.asarray(sigma, dtype=np.float64)
    return np.repeat(Q[None], len(sigma), 0), np.stack([s ** 2 * R for s in sigma])


def fitz_rhs(x0, theta):
    """f(x0) for FitzHugh-Nagumo, batched: x0 (B,2), theta (B,3) -> (B,2)"""
    a, b, c = theta[:, 0], theta[:, 1], theta[:, 2]
    V, R = x0[:, 0], x0[:, 1]
    return np.stack([c * (V - V * V * V / 3 + R), -1 / c * (V - a + b * R)], axis=1)


def fitz_problem(B, n_steps=800, t_max=40.0, sigma=0.1, seed=0, jitter=True):
    """C1/C2/C5: FitzHugh-Nagumo, nb=2, p=3 (reference README.md:101-126; theta law of SURVEY 8(d) C2)."""
    rng = np.random.default_rng(seed)
    theta0 = np.array([0.2, 0.2, 3.0]); x00 = np.array([-1.0, 1.0])
    if jitter:
        theta = np.exp(np.log(theta0) + 0.1 * rng.standard_normal((B, 3)))
        x0 = x00 + 0.05 * rng.standard_normal((B, 2))
    else:
        theta = np.tile(theta0, (B, 1)); x0 = np.tile(x00, (B, 1))
    X0 = np.zeros((B, 2, 3)); X0[:, :, 0] = x0; X0[:, :, 1] = fitz_rhs(x0, theta)
    W = np.zeros((2, 1, 3)); W[:, :, 1] = 1.0
    Q, R = _ibm(t_max / n_steps, 3, [sigma, sigma])
    return dict(model="fitzhugh_nagumo", W=W, X0=X0, theta=theta, x0=x0, Q=Q, R=R, t_min=0.0, t_max=t_max,
                n_steps=n_steps)


def fitz_obs(prob, truth_mean=None, n_obs=41, noise_var=0.005, seed=1):
    """Observations at t = 0, 1, ..., t_max (reference docs/examples/parameter.md:99-151, 420-425)."""
    rng = np.random.default_rng(seed)
    t_max, N = prob["t_max"], prob["n_steps"]
    obs_times = np.linspace(0.0, t_max, n_obs)
    if truth_mean is None:
        # smooth synthetic signal of the right scale if no solver output is supplied
        truth = np.stack([2 * np.cos(obs_times / 2), np.sin(obs_times / 2)], axis=1)
    else:
        ind = np.searchsorted(np.linspace(0.0, t_max, N + 1), obs_times)
        truth = truth_mean[ind, :, 0]
    Y = truth + np.sqrt(noise_var) * rng.standard_normal(truth.shape)
    obs_data = Y[:, :, None]
    obs_weight = np.zeros((n_obs, 2, 1, 3)); obs_weight[..., 0] = 1.0
    obs_var = np.full((n_obs, 2, 1, 1), noise_var)
    return dict(obs_data=obs_data, obs_times=obs_times, obs_weight=obs_weight, obs_var=obs_var)


def lorenz_problem(B, n_steps=4000, t_max=20.0, sigma=5e7, seed=0):
    """C3: Lorenz63, nb=3, p=3 (reference docs/examples/lorenz.md:70,101-117)."""
    rng = np.random.default_rng(seed)
    theta = np.exp(np.log(np.array([28.0, 10.0, 8.0 / 3.0])) + 0.01 * rng.standard_normal((B, 3)))
    x0 = np.tile(np.array([-12.0, -5.0, 38.0]), (B, 1))
    rho, sg, beta = theta[:, 0], theta[:, 1], theta[:, 2]
    x, y, z = x0[:, 0], x0[:, 1], x0[:, 2]
    f = np.stack([-sg * x + sg * y, rho * x - y - x * z, -beta * z + x * y], axis=1)
    X0 = np.zeros((B, 3, 3)); X0[:, :, 0] = x0; X0[:, :, 1] = f
    W = np.zeros((3, 1, 3)); W[:, :, 1] = 1.0
    Q, R = _ibm(t_max / n_steps, 3, [sigma] * 3)
    return dict(model="lorenz63", W=W, X0=X0, theta=theta, x0=x0, Q=Q, R=R, t_min=0.0, t_max=t_max,
                n_steps=n_steps)


def second_order_problem(B, n_steps=2000, t_max=10.0, sigma=0.001, seed=0):
    """C4: x'' = sin(omega t) - k x, nb=1, p=4 (reference docs/examples/higher_order.md:47-78)."""
    rng = np.random.default_rng(seed)
    theta = np.array([2.0, 1.0]) * np.exp(0.05 * rng.standard_normal((B, 2)))
    X0 = np.tile(np.array([[-1.0, 0.0, 1.0, 0.0]]), (B, 1, 1))
    X0[:, 0, 2] = theta[:, 1]
    W = np.array([[[0.0, 0.0, 1.0, 0.0]]])
    Q, R = _ibm(t_max / n_steps, 4, [sigma])
    return dict(model="second_order_sin", W=W, X0=X0, theta=theta, Q=Q, R=R, t_min=0.0, t_max=t_max,
                n_steps=n_steps)


def second_order_obs(prob, n_obs=11, noise_var=0.005, seed=1):
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, prob["t_max"], n_obs)
    exact = (-3 * np.cos(t) + 2 * np.sin(t) - np.sin(2 * t)) / 3
    obs_data = (exact + np.sqrt(noise_var) * rng.standard_normal(n_obs))[:, None, None]
    obs_weight = np.zeros((n_obs, 1, 1, 4)); obs_weight[..., 0] = 1.0
    obs_var = np.full((n_obs, 1, 1, 1), noise_var)
    return dict(obs_data=obs_data, obs_times=t, obs_weight=obs_weight, obs_var=obs_var)


def maxnorm_rel(a, b):
    """max|a-b| / max|b|: the parity metric (see tests/test_oracle_solve.py)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
