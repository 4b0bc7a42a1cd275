"""
Pins the oracle's solver / likelihood compositions (CPU only):
  * batched solver == an independent per-theta, per-block loop statement (the role of reference
    tests/test_rodeofor.py:93-121 + tests/ode_block_solve_for.py:81-235);
  * FitzHugh-Nagumo vs scipy odeint (reference tests/test_fitz.py:16-29, rel_err <= 5.0, plus a tighter bound);
  * docs' second-order ODE vs its analytic solution (reference docs/examples/higher_order.md:149-156);
  * dalton == fenrir == exact dense-Gaussian log p(Y | Z = 0) for a linear ODE under interrogate_kramer
    (the only available pin for rodeo.inference, which the reference never tests).
"""
import numpy as np
import pytest
from scipy.integrate import odeint

from oracle import rodeo_oracle as orc


def maxnorm_rel(a, b):
    """max|a-b| / max|b| -- the parity metric used throughout (two equivalent f64 statements of the recursion
    differ by ~1e-13 in this norm; element-wise relative error is meaningless for entries that pass through 0)."""
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def fitz_setup(n_steps=200, t_max=10.0, sigma=0.001, B=1, seed=None):
    mdl = orc.MODELS["fitzhugh_nagumo"]
    W, init = orc.first_order_pad(mdl, 2, 3)
    theta = np.tile(np.array([0.2, 0.2, 3.0]), (B, 1))
    x0 = np.tile(np.array([-1.0, 1.0]), (B, 1))
    if seed is not None:
        rng = np.random.default_rng(seed)
        theta = theta * np.exp(0.1 * rng.standard_normal(theta.shape))
        x0 = x0 + 0.05 * rng.standard_normal(x0.shape)
    X0 = init(x0, 0.0, theta)
    prior = orc.ibm_init(t_max / n_steps, 3, np.array([sigma] * 2))
    return mdl, W, X0, theta, x0, prior


def test_first_order_pad_matches_reference_fixture():
    # reference tests/utils.py:84-85: x0_block = [[-1, 1, 0], [1, 1/3, 0]] for theta = (.2, .2, 3)
    mdl, W, X0, *_ = fitz_setup()
    assert np.allclose(X0[0], [[-1.0, 1.0, 0.0], [1.0, 1.0 / 3.0, 0.0]], rtol=1e-15)
    assert W.shape == (2, 1, 3) and np.array_equal(W[:, 0], [[0, 1, 0], [0, 1, 0]])


def _loop_solver(mdl, W, X0, theta, t_min, t_max, N, Q, R, kind):
    """Independent un-batched statement with explicit (t, block) loops and 1-D/2-D NumPy only."""
    nb, p = X0.shape
    m = W.shape[1]
    mf = np.zeros((N + 1, nb, p)); vf = np.zeros((N + 1, nb, p, p))
    mp = np.zeros((N + 1, nb, p)); vp = np.zeros((N + 1, nb, p, p))
    mf[0] = X0; mp[0] = X0
    for t in range(N):
        for b in range(nb):
            mp[t + 1, b] = Q[b] @ mf[t, b]
            vp[t + 1, b] = Q[b] @ vf[t, b] @ Q[b].T + R[b]
        tt = t_min + (t_max - t_min) * (t + 1) / N
        f = mdl.fun(mp[t + 1][None], tt, theta[None])[0]
        J = mdl.jac(mp[t + 1][None], tt, theta[None])[0]
        for b in range(nb):
            if kind == "kramer":
                Wm = W[b] - J[b]; d = -f[b] + J[b] @ mp[t + 1, b]; V = np.zeros((m, m))
            elif kind == "schober":
                Wm = W[b]; d = -f[b]; V = np.zeros((m, m))
            else:  # rodeo
                Wm = W[b]; d = -f[b]; V = W[b] @ vp[t + 1, b] @ W[b].T
            S = Wm @ vp[t + 1, b] @ Wm.T + V
            K = np.linalg.solve(S, Wm @ vp[t + 1, b]).T
            mf[t + 1, b] = mp[t + 1, b] + K @ (0.0 - (Wm @ mp[t + 1, b] + d))
            vf[t + 1, b] = vp[t + 1, b] - K @ (Wm @ vp[t + 1, b])
    ms = np.zeros_like(mf); vs = np.zeros_like(vf)
    ms[0] = X0; ms[N] = mf[N]; vs[N] = vf[N]
    for t in range(N - 1, 0, -1):
        for b in range(nb):
            G = np.linalg.solve(vp[t + 1, b], Q[b] @ vf[t, b]).T
            ms[t, b] = mf[t, b] + G @ (ms[t + 1, b] - mp[t + 1, b])
            vs[t, b] = vf[t, b] + G @ (vs[t + 1, b] - vp[t + 1, b]) @ G.T
    return (mp, vp, mf, vf), (ms, vs)


@pytest.mark.parametrize("kind,interr", [("kramer", orc.interrogate_kramer), ("schober", orc.interrogate_schober),
                                         ("rodeo", orc.interrogate_rodeo)])
def test_batched_solver_equals_loop_solver(kind, interr):
    N, t_max = 60, 3.0
    mdl, W, X0, theta, _, prior = fitz_setup(N, t_max, sigma=0.1, B=3, seed=4)
    Q, R = prior
    filt = orc.solve_filter(mdl, W, X0, 0.0, t_max, N, interr, Q, R, theta)
    ms, vs = orc.solve_mv(mdl, W, X0, 0.0, t_max, N, interr, prior, theta)
    for i in range(3):
        lf, (lms, lvs) = _loop_solver(mdl, W, X0[i], theta[i], 0.0, t_max, N, Q, R, kind)
        for a, b in zip(filt, lf):
            assert maxnorm_rel(a[i], b) < 1e-11
        assert maxnorm_rel(ms[i], lms) < 1e-11 and maxnorm_rel(vs[i], lvs) < 1e-11
    # boundary rows (reference src/rodeo/solve.py:295-301)
    assert np.array_equal(ms[:, 0], X0) and not vs[:, 0].any()
    assert np.array_equal(ms[:, N], filt[2][:, N]) and np.array_equal(vs[:, N], filt[3][:, N])


def test_fitz_vs_odeint():
    # reference tests/test_fitz.py set-up: t in [0,10], h=.05, sigma=.001, interrogate_rodeo
    mdl, W, X0, theta, x0, prior = fitz_setup()
    tseq = np.linspace(0, 10, 201)

    def rhs(X, t, th):
        a, b, c = th
        V, R = X
        return np.array([c * (V - V * V * V / 3 + R), -1 / c * (V - a + b * R)])

    det = odeint(rhs, x0[0], tseq, args=(theta[0],))
    for interr in (orc.interrogate_rodeo, orc.interrogate_kramer):
        m, _ = orc.solve_mv(mdl, W, X0, 0.0, 10.0, 200, interr, prior, theta)
        x1, x2 = m[0, :, :, 0].ravel(), det.ravel()
        assert np.max(np.abs((x1 - x2) / (0.1 + x1))) <= 5.0          # the reference's own assertion
        assert np.max(np.abs(m[0, :, :, 0] - det)) < 0.15              # and a meaningful one (h=.05 discretisation error)
    rng = np.random.default_rng(0)
    sim = orc.solve_sim(mdl, W, X0, 0.0, 10.0, 200, orc.interrogate_rodeo, prior, theta,
                        z_smooth=rng.standard_normal((1, 201, 2, 3)))
    assert np.max(np.abs(sim[0, :, :, 0] - det)) < 0.15


def so_setup(N, sigma, B=1, seed=None, t_max=10.0):
    mdl = orc.MODELS["second_order_sin"]
    W = np.array([[[0.0, 0.0, 1.0, 0.0]]])
    theta = np.tile(np.array([2.0, 1.0]), (B, 1))
    if seed is not None:
        theta = theta * np.exp(0.05 * np.random.default_rng(seed).standard_normal(theta.shape))
    X0 = np.tile(np.array([[-1.0, 0.0, 1.0, 0.0]]), (B, 1, 1))
    X0[:, 0, 2] = theta[:, 1]                    # x''(0) = sin(0) - k x(0) = k
    prior = orc.ibm_init(t_max / N, 4, np.array([sigma]))
    return mdl, W, X0, theta, prior


def test_second_order_vs_analytic():
    N = 400
    mdl, W, X0, theta, prior = so_setup(N, 0.001)
    m, _ = orc.solve_mv(mdl, W, X0, 0.0, 10.0, N, orc.interrogate_kramer, prior, theta)
    t = np.linspace(0, 10, N + 1)
    exact_x = (-3 * np.cos(t) + 2 * np.sin(t) - np.sin(2 * t)) / 3
    exact_x1 = (-2 * np.cos(2 * t) + 3 * np.sin(t) + 2 * np.cos(t)) / 3
    assert np.max(np.abs(m[0, :, 0, 0] - exact_x)) < 2e-3
    assert np.max(np.abs(m[0, :, 0, 1] - exact_x1)) < 2e-3


def _dense_logpdf(x, mean, cov):
    L = np.linalg.cholesky(cov)
    r = np.linalg.solve(L, x - mean)
    return -0.5 * r @ r - np.log(np.diag(L)).sum() - 0.5 * len(x) * np.log(2 * np.pi)


def test_dalton_equals_fenrir_equals_exact_for_linear_ode():
    # linear ODE + kramer => W~ = W - J and d = -sin(w t) are state independent: exactly linear-Gaussian.
    N, t_max, sigma = 24, 6.0, 1.0
    mdl, W, X0, theta, prior = so_setup(N, sigma, t_max=t_max)
    Q, R = prior[0][0], prior[1][0]
    p = 4
    obs_times = np.array([0.0, 1.0, 2.5, 4.0, 6.0])
    n_obs = len(obs_times)
    rng = np.random.default_rng(7)
    obs_data = rng.standard_normal((n_obs, 1, 1))
    obs_weight = np.zeros((n_obs, 1, 1, p)); obs_weight[..., 0] = 1.0
    obs_var = np.full((n_obs, 1, 1, 1), 0.3)
    ll_d = orc.dalton(mdl, W, X0, 0.0, t_max, N, orc.interrogate_kramer, prior, theta,
                      obs_data, obs_times, obs_weight, obs_var)[0]
    ll_f = orc.fenrir(mdl, W, X0, 0.0, t_max, N, orc.interrogate_kramer, prior, theta,
                      obs_data, obs_times, obs_weight, obs_var)[0]

    # dense joint of X_{1:N}; Z_n = Wt X_n + d_n ; Y_i = D X_{n(i)} + e
    om, k = theta[0]
    Wt = np.array([k, 0.0, 1.0, 0.0])                               # W - J, J = [-k,0,0,0]
    mean_x = np.zeros((N + 1, p)); mean_x[0] = X0[0, 0]
    Phi = np.zeros((N + 1, N + 1, p, p))                            # cov(X_a, X_b)
    var = np.zeros((p, p))
    covs = [var]
    for n in range(1, N + 1):
        mean_x[n] = Q @ mean_x[n - 1]
        var = Q @ var @ Q.T + R
        covs.append(var)
    for a in range(N + 1):
        Phi[a, a] = covs[a]
        M = covs[a]
        for b in range(a + 1, N + 1):
            M = M @ Q.T
            Phi[a, b] = M; Phi[b, a] = M.T
    tt = np.array([t_max * n / N for n in range(N + 1)])
    Hz = np.zeros((N, (N + 1) * p)); dz = np.zeros(N)
    for n in range(1, N + 1):
        Hz[n - 1, n * p:(n + 1) * p] = Wt
        dz[n - 1] = -np.sin(om * tt[n])
    ind = orc.obs_index(0.0, t_max, N, obs_times)
    Hy = np.zeros((n_obs, (N + 1) * p))
    for i, n in enumerate(ind):
        Hy[i, n * p] = 1.0
    big = Phi.transpose(0, 2, 1, 3).reshape((N + 1) * p, (N + 1) * p)
    mx = mean_x.ravel()
    H = np.vstack([Hz, Hy])
    off = np.concatenate([dz, np.zeros(n_obs)])
    noise = np.diag(np.concatenate([np.zeros(N), np.full(n_obs, 0.3)]))
    mean_zy = H @ mx + off
    cov_zy = H @ big @ H.T + noise
    target = np.concatenate([np.zeros(N), obs_data.ravel()])
    exact = _dense_logpdf(target, mean_zy, cov_zy) - _dense_logpdf(np.zeros(N), mean_zy[:N], cov_zy[:N, :N])
    assert abs(ll_d - exact) < 1e-7 * max(1.0, abs(exact)), (ll_d, exact)
    assert abs(ll_f - exact) < 1e-7 * max(1.0, abs(exact)), (ll_f, exact)


def test_dalton_obs_pointer_semantics():
    # obs_ind[0] == 0 adds log N(y0; D x0, Omega) and starts the pointer at 1 (dalton.py:207-215);
    # an observation time strictly inside a grid cell maps to the next grid point (left insertion).
    assert list(orc.obs_index(0.0, 4.0, 8, [0.0, 0.9, 1.0, 4.0])) == [0, 2, 2, 8]
    N, t_max = 16, 4.0
    mdl, W, X0, theta, _, prior = fitz_setup(N, t_max, sigma=0.5, B=2, seed=1)
    rng = np.random.default_rng(0)
    obs_times = np.array([0.0, 1.0, 2.0, 4.0])
    obs_data = rng.standard_normal((4, 2, 1))
    D = np.zeros((4, 2, 1, 3)); D[..., 0] = 1.0
    Om = np.full((4, 2, 1, 1), 0.05)
    full = orc.dalton(mdl, W, X0, 0.0, t_max, N, orc.interrogate_kramer, prior, theta, obs_data, obs_times, D, Om)
    tail = orc.dalton(mdl, W, X0, 0.0, t_max, N, orc.interrogate_kramer, prior, theta,
                      obs_data[1:], obs_times[1:], D[1:], Om[1:])
    y0 = np.array([sum(orc.multivariate_normal_logpdf(obs_data[0, b], D[0, b] @ X0[k, b], Om[0, b]) for b in range(2))
                   for k in range(2)])
    assert np.allclose(full, tail + y0, rtol=1e-12)


def test_fenrir_and_dalton_data_adaptive_posteriors_agree_for_a_linear_ode():
    # both are the exact posterior p(X | Z = 0, Y) of a linear-Gaussian model: pins the stacks / index conventions of
    # fenrir._backward + _smooth_mv (fenrir.py:86-401) against dalton._solve_filter + smoother (dalton.py:242-460)
    mdl, W, X0, theta, prior = so_setup(60, 1.0, B=3, seed=2, t_max=3.0)
    obs_times = np.array([0.0, 1.0, 2.0, 3.0])
    rng = np.random.default_rng(5)
    y = rng.standard_normal((4, 1, 1))
    D = np.zeros((4, 1, 1, 4)); D[..., 0] = 1.0
    Om = np.full((4, 1, 1, 1), 0.05)
    a = orc.fenrir_solve_mv(mdl, W, X0, 0.0, 3.0, 60, orc.interrogate_kramer, prior, theta, y, obs_times, D, Om)
    b = orc.dalton_solve_mv(mdl, W, X0, 0.0, 3.0, 60, orc.interrogate_kramer, prior, theta, y, obs_times, D, Om)
    assert maxnorm_rel(a[0], b[0]) < 1e-9 and maxnorm_rel(a[1], b[1]) < 1e-9
