"""GPU: the reference's own known-answer tests of the Kalman primitives (reference tests/test_standard.py:18-200,
tests/utils.py:117-215), restated against the CUDA primitives: every predicted / filtered / smoothed moment of a random
state-space model must equal brute-force conditioning of its dense joint Gaussian.  The primitives called here are the
__device__ functions the fused solver kernels inline, exposed one-thread-per-problem through the C ABI."""
import numpy as np
import pytest
import scipy.stats

import gm_bruteforce as gm

pytestmark = pytest.mark.gpu


def ref_rel_err(x1, x2):
    """reference tests/utils.py:11-18"""
    x1 = np.ravel(x1) * 1.0; x2 = np.ravel(x2) * 1.0
    return np.max(np.abs((x1 - x2) / (0.1 + x1)))


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ktv():
    import rodeo_b200
    return rodeo_b200.kalmantv.standard


@pytest.mark.parametrize("seed", range(10))
def test_primitives_match_bruteforce_conditioning(ktv, seed):
    rng = np.random.default_rng(seed)
    ssm = gm.random_ssm(rng)                       # n_meas in 1..3, n_state = n_meas + 1..4, as the reference draws them
    mean, cov = gm.joint_gaussian(ssm)
    T, s = ssm["n_tot"], ssm["n_state"]
    filt, pred = [], []
    m_p, v_p = ssm["mean_state"][0], ssm["var_state"][0]
    for n in range(T):
        if n > 0:
            m_p, v_p = map(_np, ktv.predict(mean_state_past=filt[-1][0], var_state_past=filt[-1][1],
                                            mean_state=ssm["mean_state"][n], wgt_state=ssm["wgt_state"][n - 1],
                                            var_state=ssm["var_state"][n]))
        pred.append((m_p, v_p))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n - 1)
        assert ref_rel_err(bm, m_p) < 5e-8 and ref_rel_err(bv, v_p) < 5e-8          # the reference's assertAlmostEqual
        m_f, v_f = map(_np, ktv.update(mean_state_pred=m_p, var_state_pred=v_p, x_meas=ssm["x_meas"][n],
                                       mean_meas=ssm["mean_meas"][n], wgt_meas=ssm["wgt_meas"][n],
                                       var_meas=ssm["var_meas"][n]))
        filt.append((m_f, v_f))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n)
        assert ref_rel_err(bm, m_f) < 5e-8 and ref_rel_err(bv, v_f) < 5e-8
        assert np.max(np.abs(bm - m_f)) < 1e-9 * max(1, np.abs(bm).max())
        fm, fv = map(_np, ktv.forecast(mean_state_pred=m_p, var_state_pred=v_p, mean_meas=ssm["mean_meas"][n],
                                       wgt_meas=ssm["wgt_meas"][n], var_meas=ssm["var_meas"][n]))
        W = ssm["wgt_meas"][n]
        assert np.allclose(fm, W @ m_p + ssm["mean_meas"][n], rtol=1e-12, atol=1e-13)
        assert np.allclose(fv, W @ v_p @ W.T + ssm["var_meas"][n], rtol=1e-11, atol=1e-13)
    ms, vs = filt[T - 1]
    for n in range(T - 2, -1, -1):
        ms, vs = map(_np, ktv.smooth_mv(mean_state_next=ms, var_state_next=vs, mean_state_filt=filt[n][0],
                                        var_state_filt=filt[n][1], mean_state_pred=pred[n + 1][0],
                                        var_state_pred=pred[n + 1][1], wgt_state=ssm["wgt_state"][n]))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, T - 1)
        assert ref_rel_err(bm, ms) < 5e-8 and ref_rel_err(bv, vs) < 5e-8
    x_next = rng.standard_normal(s)
    for n in range(T - 1):
        mj, vj = gm.theta_mn(ssm, mean, cov, [n, n + 1], T - 1)
        cm, cv = gm.condition(mj, vj, np.arange(s), np.arange(s, 2 * s), x_next)
        m_sim, v_sim = map(_np, ktv.smooth_sim(x_state_next=x_next, mean_state_filt=filt[n][0],
                                               var_state_filt=filt[n][1], mean_state_pred=pred[n + 1][0],
                                               var_state_pred=pred[n + 1][1], wgt_state=ssm["wgt_state"][n]))
        assert ref_rel_err(cm, m_sim) < 5e-8 and ref_rel_err(cv, v_sim) < 5e-8
        A, b, C = map(_np, ktv.smooth_cond(mean_state_filt=filt[n][0], var_state_filt=filt[n][1],
                                           mean_state_pred=pred[n + 1][0], var_state_pred=pred[n + 1][1],
                                           wgt_state=ssm["wgt_state"][n]))
        assert np.allclose(A @ x_next + b, cm, rtol=1e-9, atol=1e-12) and np.allclose(C, cv, rtol=1e-9, atol=1e-12)


def test_primitives_are_batched(ktv):
    rng = np.random.default_rng(0)
    ssms = [gm.random_ssm(rng, n_meas=2, n_state=4) for _ in range(5)]
    stack = lambda k, i: np.stack([s[k][i] for s in ssms])
    mp, vp = ktv.predict(stack("mean_state", 0), stack("var_state", 0), stack("mean_state", 1), stack("wgt_state", 0),
                         stack("var_state", 1))
    assert mp.shape == (5, 4) and vp.shape == (5, 4, 4)
    one = ktv.predict(ssms[3]["mean_state"][0], ssms[3]["var_state"][0], ssms[3]["mean_state"][1],
                      ssms[3]["wgt_state"][0], ssms[3]["var_state"][1])
    assert np.array_equal(_np(mp[3]), _np(one[0])) and np.array_equal(_np(vp[3]), _np(one[1]))


def test_device_logpdf_matches_scipy_and_cutoff():
    import rodeo_b200
    lp = rodeo_b200.utils.multivariate_normal_logpdf
    rng = np.random.default_rng(1)
    for n in (1, 2, 3):
        a = rng.standard_normal((64, n, n))
        cov = a @ np.swapaxes(a, -1, -2) + 0.3 * np.eye(n)
        mu, x = rng.standard_normal((64, n)), rng.standard_normal((64, n))
        got = _np(lp(x, mu, cov))
        want = np.array([scipy.stats.multivariate_normal(mu[k], cov[k]).logpdf(x[k]) for k in range(64)])
        assert np.max(np.abs(got - want) / np.maximum(1, np.abs(want))) < 1e-12
    # |w| <= 1e-8 contributes nothing (reference src/rodeo/utils.py:74)
    assert float(lp(np.array([3.0]), np.array([0.0]), np.array([[0.9e-8]]))) == 0.0
    got = float(lp(np.array([5.0, 1.0]), np.zeros(2), np.diag([1e-9, 2.0])))
    assert np.isclose(got, -0.5 * (0.5 + np.log(2.0)) - 0.5 * np.log(2 * np.pi), rtol=1e-13)
    got = float(lp(np.array([5.0, 1.0, 0.3]), np.zeros(3), np.diag([1e-9, 2.0, 4.0])))
    assert np.isclose(got, -0.5 * (0.5 + np.log(2.0) + 0.09 / 4 + np.log(4.0)) - np.log(2 * np.pi), rtol=1e-13)


# ---- square-root family (reference tests/test_square_root.py:19-168) ---------------------------------------------------
@pytest.mark.parametrize("seed", range(10))
def test_square_root_primitives_match_bruteforce_conditioning(seed):
    """The reference's known-answer tests of rodeo.kalmantv.square_root, restated against the CUDA primitives: with
    Cholesky factors in, L L^T of every factor out equals brute-force conditioning of the dense joint Gaussian (the
    reference compares L L^T too, and only to 2 decimal places for the smooth_sim / smooth_cond variances)."""
    import rodeo_b200
    sq = rodeo_b200.kalmantv.square_root
    chol = np.linalg.cholesky
    sqr = lambda L: _np(L) @ _np(L).T
    rng = np.random.default_rng(100 + seed)
    ssm = gm.random_ssm(rng)
    mean, cov = gm.joint_gaussian(ssm)
    T, s = ssm["n_tot"], ssm["n_state"]
    filt, pred = [], []
    m_p, L_p = ssm["mean_state"][0], chol(ssm["var_state"][0])
    for n in range(T):
        if n > 0:
            m_p, L_p = sq.predict(mean_state_past=filt[-1][0], var_state_past=filt[-1][1],
                                  mean_state=ssm["mean_state"][n], wgt_state=ssm["wgt_state"][n - 1],
                                  var_state=chol(ssm["var_state"][n]))
            m_p, L_p = _np(m_p), _np(L_p)
            assert np.array_equal(np.triu(L_p, 1), np.zeros_like(L_p))
        pred.append((m_p, L_p))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n - 1)
        assert ref_rel_err(bm, m_p) < 5e-8 and ref_rel_err(bv, L_p @ L_p.T) < 5e-8
        m_f, L_f = map(_np, sq.update(mean_state_pred=m_p, var_state_pred=L_p, x_meas=ssm["x_meas"][n],
                                      mean_meas=ssm["mean_meas"][n], wgt_meas=ssm["wgt_meas"][n],
                                      var_meas=chol(ssm["var_meas"][n])))
        filt.append((m_f, L_f))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n)
        assert ref_rel_err(bm, m_f) < 5e-8 and ref_rel_err(bv, L_f @ L_f.T) < 5e-8
        fm, fv = map(_np, sq.forecast(mean_state_pred=m_p, var_state_pred=L_p, mean_meas=ssm["mean_meas"][n],
                                      wgt_meas=ssm["wgt_meas"][n], var_meas=chol(ssm["var_meas"][n])))
        W = ssm["wgt_meas"][n]
        assert np.allclose(fm, W @ m_p + ssm["mean_meas"][n], rtol=1e-12, atol=1e-13)
        assert np.allclose(fv, W @ (L_p @ L_p.T) @ W.T + ssm["var_meas"][n], rtol=1e-10, atol=1e-12)
    ms, Ls = filt[T - 1]
    for n in range(T - 2, -1, -1):
        ms, Ls = map(_np, sq.smooth_mv(mean_state_next=ms, var_state_next=Ls, mean_state_filt=filt[n][0],
                                       var_state_filt=filt[n][1], mean_state_pred=pred[n + 1][0],
                                       var_state_pred=pred[n + 1][1], wgt_state=ssm["wgt_state"][n],
                                       var_state=chol(ssm["var_state"][n + 1])))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, T - 1)
        assert ref_rel_err(bm, ms) < 5e-8 and ref_rel_err(bv, Ls @ Ls.T) < 5e-8
    x_next = rng.standard_normal(s)
    for n in range(T - 1):
        mj, vj = gm.theta_mn(ssm, mean, cov, [n, n + 1], T - 1)
        cm, cv = gm.condition(mj, vj, np.arange(s), np.arange(s, 2 * s), x_next)
        kw = dict(mean_state_filt=filt[n][0], var_state_filt=filt[n][1], mean_state_pred=pred[n + 1][0],
                  var_state_pred=pred[n + 1][1], wgt_state=ssm["wgt_state"][n], var_state=chol(ssm["var_state"][n + 1]))
        m_sim, L_sim = sq.smooth_sim(x_state_next=x_next, **kw)
        assert ref_rel_err(cm, _np(m_sim)) < 5e-8 and ref_rel_err(cv, sqr(L_sim)) < 5e-8
        A, b, Lc = sq.smooth_cond(**kw)
        assert np.allclose(_np(A) @ x_next + _np(b), cm, rtol=1e-8, atol=1e-11)
        assert np.allclose(sqr(Lc), cv, rtol=1e-8, atol=1e-11)
        # and the covariance-form primitives agree with the squared factors
        std = rodeo_b200.kalmantv.standard
        m2, v2 = std.smooth_sim(x_state_next=x_next, mean_state_filt=filt[n][0], var_state_filt=filt[n][1] @ filt[n][1].T,
                                mean_state_pred=pred[n + 1][0], var_state_pred=pred[n + 1][1] @ pred[n + 1][1].T,
                                wgt_state=ssm["wgt_state"][n])
        assert np.allclose(_np(m2), _np(m_sim), rtol=1e-9, atol=1e-12) and np.allclose(_np(v2), sqr(L_sim), rtol=1e-8, atol=1e-11)
