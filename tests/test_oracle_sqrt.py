"""Pins the oracle's square-root Kalman family (CPU): L L^T must equal the brute-force covariances (the reference's
tests/test_square_root.py:11-168 compares exactly that), and the square-root solver must agree with the covariance-form
solver where both are well conditioned."""
import numpy as np
import pytest

import gm_bruteforce as gm
import problems as P
from oracle import rodeo_oracle as orc


def sq(L):
    return L @ np.swapaxes(L, -1, -2)


def test_add_sqrt_squares_to_the_sum():
    # reference tests/test_add_sqrt.py:6-20
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((2, 2)), rng.standard_normal((2, 2))
    A, B = A @ A.T, B @ B.T
    S = orc.add_sqrt(np.linalg.cholesky(A), np.linalg.cholesky(B))
    assert np.allclose(sq(S), A + B, rtol=1e-12) and np.allclose(np.triu(S, 1), 0)


@pytest.mark.parametrize("seed", range(4))
def test_sqrt_filter_and_smoother_match_bruteforce(seed):
    rng = np.random.default_rng(seed)
    ssm = gm.random_ssm(rng)
    mean, cov = gm.joint_gaussian(ssm)
    T = ssm["n_tot"]
    chol = lambda a: np.linalg.cholesky(a)
    filt, pred = [], []
    m_p, l_p = ssm["mean_state"][0], chol(ssm["var_state"][0])
    for n in range(T):
        if n > 0:
            m_p, l_p = orc.sqrt_predict(filt[-1][0], filt[-1][1], ssm["mean_state"][n], ssm["wgt_state"][n - 1],
                                        chol(ssm["var_state"][n]))
        pred.append((m_p, l_p))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n - 1)
        assert np.allclose(m_p, bm, rtol=1e-9, atol=1e-12) and np.allclose(sq(l_p), bv, rtol=1e-8, atol=1e-11)
        m_f, l_f = orc.sqrt_update(m_p, l_p, ssm["x_meas"][n], ssm["mean_meas"][n], ssm["wgt_meas"][n],
                                   chol(ssm["var_meas"][n]))
        filt.append((m_f, l_f))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, n)
        assert np.allclose(m_f, bm, rtol=1e-9, atol=1e-12) and np.allclose(sq(l_f), bv, rtol=1e-8, atol=1e-11)
    ms, ls = filt[T - 1]
    for n in range(T - 2, -1, -1):
        ms, ls = orc.sqrt_smooth_mv(ms, ls, filt[n][0], filt[n][1], pred[n + 1][0], pred[n + 1][1],
                                    ssm["wgt_state"][n], chol(ssm["var_state"][n + 1]))
        bm, bv = gm.theta_mn(ssm, mean, cov, n, T - 1)
        assert np.allclose(ms, bm, rtol=1e-9, atol=1e-12) and np.allclose(sq(ls), bv, rtol=1e-8, atol=1e-11)


def test_sqrt_solver_agrees_with_covariance_solver():
    pr = P.fitz_problem(6, n_steps=120, t_max=6.0, seed=2)
    Rh = np.linalg.cholesky(pr["R"])
    mdl = orc.MODELS["fitzhugh_nagumo"]
    m1, v1 = orc.solve_mv(mdl, pr["W"], pr["X0"], 0.0, 6.0, 120, orc.interrogate_kramer, (pr["Q"], pr["R"]), pr["theta"])
    m2, l2 = orc.solve_mv_sqrt(mdl, pr["W"], pr["X0"], 0.0, 6.0, 120, orc.interrogate_kramer, (pr["Q"], Rh), pr["theta"])
    assert P.maxnorm_rel(m2, m1) < 1e-9 and P.maxnorm_rel(sq(l2), v1) < 1e-8
    assert not l2[:, 0].any() and np.allclose(np.triu(l2, 1), 0)
