"""
Brute-force Gaussian conditioning for a small linear state-space model.

Restates, in NumPy, the known-answer procedure the reference's own Kalman tests use
(reference tests/utils.py:24-63,117-154 `kalman_theta` / `kalman_setup`, tests/gauss_markov.py:30-125):
build the dense joint Gaussian of (x_0, y_0, ..., x_T, y_T) for

    x_0 = c_0 + R_0^{1/2} e_0,   x_n = c_n + Q_n x_{n-1} + R_n^{1/2} e_n,   y_n = d_n + W_n x_n + V_n^{1/2} h_n

and read E[x_m | y_{0:n}], var(x_m | y_{0:n}) off it by direct conditioning.  Test infrastructure only.
"""
import numpy as np


def random_ssm(rng, n_tot=3, n_meas=None, n_state=None):
    """Random model with the size ranges of reference tests/utils.py:117-147."""
    if n_meas is None:
        n_meas = int(rng.integers(1, 4))
    if n_state is None:
        n_state = n_meas + int(rng.integers(1, 5))
    sq = lambda a: a @ np.swapaxes(a, -1, -2)
    return dict(
        n_tot=n_tot, n_meas=n_meas, n_state=n_state,
        mean_state=rng.standard_normal((n_tot, n_state)),
        var_state=sq(rng.standard_normal((n_tot, n_state, n_state))),
        wgt_state=0.01 * rng.standard_normal((n_tot - 1, n_state, n_state)),
        mean_meas=rng.standard_normal((n_tot, n_meas)),
        var_meas=sq(rng.standard_normal((n_tot, n_meas, n_meas))),
        wgt_meas=rng.standard_normal((n_tot, n_meas, n_state)),
        x_meas=rng.standard_normal((n_tot, n_meas)),
    )


def joint_gaussian(ssm):
    """Mean (T*(s+m),) and covariance of the stacked vector [x_0..x_{T-1}, y_0..y_{T-1}]."""
    T, s, m = ssm["n_tot"], ssm["n_state"], ssm["n_meas"]
    # X = c + Qshift X + noise  =>  X = (I - Qshift)^{-1} (c + noise)
    A = np.eye(T * s)
    for n in range(1, T):
        A[n * s:(n + 1) * s, (n - 1) * s:n * s] = -ssm["wgt_state"][n - 1]
    Ainv = np.linalg.inv(A)
    Rbig = np.zeros((T * s, T * s))
    Wbig = np.zeros((T * m, T * s))
    Vbig = np.zeros((T * m, T * m))
    for n in range(T):
        Rbig[n * s:(n + 1) * s, n * s:(n + 1) * s] = ssm["var_state"][n]
        Wbig[n * m:(n + 1) * m, n * s:(n + 1) * s] = ssm["wgt_meas"][n]
        Vbig[n * m:(n + 1) * m, n * m:(n + 1) * m] = ssm["var_meas"][n]
    mx = Ainv @ ssm["mean_state"].ravel()
    Sxx = Ainv @ Rbig @ Ainv.T
    my = ssm["mean_meas"].ravel() + Wbig @ mx
    Sxy = Sxx @ Wbig.T
    Syy = Wbig @ Sxx @ Wbig.T + Vbig
    mean = np.concatenate([mx, my])
    cov = np.block([[Sxx, Sxy], [Sxy.T, Syy]])
    return mean, cov


def condition(mean, cov, ikeep, icond, vals):
    """Moments of v[ikeep] | v[icond] = vals."""
    S22 = cov[np.ix_(icond, icond)]
    S12 = cov[np.ix_(ikeep, icond)]
    A = np.linalg.solve(S22, S12.T).T
    mu = mean[ikeep] + A @ (vals - mean[icond])
    V = cov[np.ix_(ikeep, ikeep)] - A @ S12.T
    return mu, V


def theta_mn(ssm, mean, cov, m_idx, n_obs_upto):
    """E / var of x_{m_idx} (int or list) given y_0..y_{n_obs_upto} (inclusive; -1 = no data)."""
    T, s, m = ssm["n_tot"], ssm["n_state"], ssm["n_meas"]
    m_idx = np.atleast_1d(m_idx)
    ikeep = np.concatenate([np.arange(k * s, (k + 1) * s) for k in m_idx])
    icond = T * s + np.arange(0, (n_obs_upto + 1) * m)
    vals = ssm["x_meas"][:n_obs_upto + 1].ravel()
    if len(icond) == 0:
        return mean[ikeep], cov[np.ix_(ikeep, ikeep)]
    return condition(mean, cov, ikeep, icond, vals)
