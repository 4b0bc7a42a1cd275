"""
Pins the oracle's Kalman primitives the way the reference pins its own (reference tests/test_standard.py:18-200):
every filtered / predicted / smoothed moment must equal brute-force conditioning of the dense joint Gaussian.
CPU only.
"""
import numpy as np
import pytest
import scipy.stats

import gm_bruteforce as gm
from oracle import rodeo_oracle as orc


def ref_rel_err(x1, x2):
    """reference tests/utils.py:11-18 (denominator 0.1 + x1, kept verbatim for the reference's assertion)"""
    x1 = np.ravel(x1) * 1.0
    x2 = np.ravel(x2) * 1.0
    return np.max(np.abs((x1 - x2) / (0.1 + x1)))


def true_rel_err(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300)


@pytest.mark.parametrize("seed", range(8))
def test_filter_and_smoothers_match_bruteforce(seed):
    rng = np.random.default_rng(seed)
    ssm = gm.random_ssm(rng)
    mean, cov = gm.joint_gaussian(ssm)
    T, s = ssm["n_tot"], ssm["n_state"]

    # forward filter with the oracle
    filt, pred = [], []
    m_p, v_p = ssm["mean_state"][0], ssm["var_state"][0]        # theta_{0|-1}
    for n in range(T):
        if n > 0:
            m_p, v_p = orc.predict(filt[-1][0], filt[-1][1], ssm["mean_state"][n],
                                   ssm["wgt_state"][n - 1], ssm["var_state"][n])
        pred.append((m_p, v_p))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n - 1)
        assert ref_rel_err(bm, m_p) < 5e-8 and ref_rel_err(bv, v_p) < 5e-8
        m_f, v_f = orc.update(m_p, v_p, ssm["x_meas"][n], ssm["mean_meas"][n], ssm["wgt_meas"][n],
                              ssm["var_meas"][n])
        filt.append((m_f, v_f))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n)
        assert ref_rel_err(bm, m_f) < 5e-8 and ref_rel_err(bv, v_f) < 5e-8
        assert true_rel_err(bm, m_f) < 1e-9 and true_rel_err(bv, v_f) < 1e-9
        # forecast = predictive moments of y_n | y_{0:n-1}
        fm, fv = orc.forecast(m_p, v_p, ssm["mean_meas"][n], ssm["wgt_meas"][n], ssm["var_meas"][n])
        W = ssm["wgt_meas"][n]
        assert np.allclose(fm, W @ m_p + ssm["mean_meas"][n]) and np.allclose(fv, W @ v_p @ W.T + ssm["var_meas"][n])

    # backward mean/variance smoother
    ms, vs = filt[T - 1]
    for n in range(T - 2, -1, -1):
        ms, vs = orc.smooth_mv(ms, vs, filt[n][0], filt[n][1], pred[n + 1][0], pred[n + 1][1],
                               ssm["wgt_state"][n])
        bm, bv = gm.theta_mn(ssm, mean, cov, n, T - 1)
        assert ref_rel_err(bm, ms) < 5e-8 and ref_rel_err(bv, vs) < 5e-8
        assert true_rel_err(bm, ms) < 1e-9 and true_rel_err(bv, vs) < 1e-9

    # sampling smoother: x_n | x_{n+1}, y_{0:T-1}  (== x_n | x_{n+1}, y_{0:n})
    x_next = rng.standard_normal(s)
    for n in range(T - 1):
        mj, vj = gm.theta_mn(ssm, mean, cov, [n, n + 1], T - 1)
        cm, cv = gm.condition(mj, vj, np.arange(s), np.arange(s, 2 * s), x_next)
        m_sim, v_sim = orc.smooth_sim(x_next, filt[n][0], filt[n][1], pred[n + 1][0], pred[n + 1][1],
                                      ssm["wgt_state"][n])
        assert ref_rel_err(cm, m_sim) < 5e-8 and ref_rel_err(cv, v_sim) < 5e-8
        A, b, C = orc.smooth_cond(filt[n][0], filt[n][1], pred[n + 1][0], pred[n + 1][1], ssm["wgt_state"][n])
        assert np.allclose(A @ x_next + b, cm, rtol=1e-9, atol=1e-12)
        assert np.allclose(C, cv, rtol=1e-9, atol=1e-12)


def test_primitives_broadcast_over_leading_axes():
    rng = np.random.default_rng(0)
    ssm = gm.random_ssm(rng, n_meas=2, n_state=4)
    m0, v0 = ssm["mean_state"][0], ssm["var_state"][0]
    one = orc.update(m0, v0, ssm["x_meas"][0], ssm["mean_meas"][0], ssm["wgt_meas"][0], ssm["var_meas"][0])
    tile = lambda a: np.broadcast_to(a, (5, 3) + a.shape).copy()
    many = orc.update(tile(m0), tile(v0), tile(ssm["x_meas"][0]), tile(ssm["mean_meas"][0]),
                      tile(ssm["wgt_meas"][0]), tile(ssm["var_meas"][0]))
    assert np.array_equal(many[0][4, 2], one[0]) and np.array_equal(many[1][4, 2], one[1])


def test_logpdf_matches_scipy_when_well_conditioned():
    rng = np.random.default_rng(1)
    for p in (1, 2, 3):
        a = rng.standard_normal((p, p))
        cov = a @ a.T + 0.5 * np.eye(p)
        mu, x = rng.standard_normal(p), rng.standard_normal(p)
        got = orc.multivariate_normal_logpdf(x, mu, cov)
        want = scipy.stats.multivariate_normal(mu, cov).logpdf(x)
        assert abs(got - want) < 1e-12 * max(1.0, abs(want))


def test_logpdf_eigenvalue_cutoff_is_absolute_1e8():
    # reference src/rodeo/utils.py:74: eigenvalues with |w| <= 1e-8 contribute nothing at all
    assert orc.multivariate_normal_logpdf(np.array([3.0]), np.array([0.0]), np.array([[0.9e-8]])) == 0.0
    kept = orc.multivariate_normal_logpdf(np.array([1e-4]), np.array([0.0]), np.array([[1.1e-8]]))
    assert np.isclose(kept, -0.5 * (1e-8 / 1.1e-8 + np.log(1.1e-8)) - 0.5 * np.log(2 * np.pi), rtol=1e-14)
    # 2x2 with one dropped direction
    cov = np.diag([1e-9, 2.0])
    got = orc.multivariate_normal_logpdf(np.array([5.0, 1.0]), np.zeros(2), cov)
    assert np.isclose(got, -0.5 * (0.5 + np.log(2.0)) - 0.5 * np.log(2 * np.pi), rtol=1e-14)


def test_ibm_prior_closed_form():
    # Q = exp(F dt), R = int_0^dt exp(F s) L L' exp(F s)' ds for the q-times integrated Brownian motion
    import scipy.linalg
    dt, q, sigma = 0.05, 2, np.array([0.1, 0.3])
    Q, R = orc.ibm_init(dt, q + 1, sigma)
    F = np.diag(np.ones(q), 1)
    assert np.allclose(Q[0], scipy.linalg.expm(F * dt), rtol=1e-14, atol=0)
    assert np.array_equal(np.tril(Q[0], -1), np.zeros((q + 1, q + 1))) and np.all(np.diag(Q[0]) == 1.0)
    for b in range(2):
        L = np.zeros((q + 1, 1)); L[q, 0] = sigma[b]
        # Van Loan: exp([[-F, LL'],[0, F']] dt) = [[., G],[0, Qt']],  R = Qt G
        Mx = np.block([[-F, L @ L.T], [np.zeros_like(F), F.T]]) * dt
        E = scipy.linalg.expm(Mx)
        Rvl = E[q + 1:, q + 1:].T @ E[:q + 1, q + 1:]
        assert np.allclose(R[b], Rvl, rtol=1e-10, atol=0)


@pytest.mark.parametrize("name", sorted(orc.MODELS))
def test_model_block_jacobians_by_complex_step(name):
    mdl = orc.MODELS[name]
    rng = np.random.default_rng(3)
    B, nb, p = 4, mdl.n_block, mdl.n_bstate
    X = rng.uniform(0.5, 1.5, (B, nb, p))
    th = rng.uniform(0.5, 1.5, (B, mdl.n_theta))
    t = 0.37
    J = mdl.jac(X, t, th)
    h = 1e-6
    for b in range(nb):
        for j in range(p):
            Xp, Xm = X.copy(), X.copy()
            Xp[:, b, j] += h; Xm[:, b, j] -= h
            fd = (mdl.fun(Xp, t, th)[:, b, :] - mdl.fun(Xm, t, th)[:, b, :]) / (2 * h)
            assert np.allclose(J[:, b, :, j], fd, rtol=1e-6, atol=1e-7), (name, b, j)
