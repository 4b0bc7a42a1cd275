"""CPU: properties of the REFERENCE's algorithm that decide what is (not) built -- restated with NumPy / LAPACK."""
import numpy as np
import scipy.linalg as sl


def test_reference_daltonng_innovation_matrix_is_exactly_singular():
    """rodeo.inference.daltonng (src/rodeo/inference/dalton.py:614-633) turns a non-Gaussian observation model into
    pseudo-observations: obs_var[b] = -pinv(Hessian[b, :, b]) (a p x p matrix), obs_weight = where(obs_var != 0, 1, 0)
    (p x p), and stacks ALL p rows under the ODE row.  For an observation model that reads one state component -- the
    docs' own example, Poisson counts with rate exp(b0 + b1 X[:, 0]), docs/examples/parameter.md:545-557 -- the Hessian
    block is diag(h, 0, ..., 0): p - 1 of the stacked rows have zero weight AND zero variance, so the innovation matrix
    W S W^T + V the update hands to jnp.linalg.solve (src/rodeo/utils.py:119; LAPACK getrf / getrs) has p - 1 zero rows
    and columns.  getrf reports an exactly zero pivot and getrs returns NaN: the reference's own result is NaN, there is
    nothing to match, and rodeo_b200 does not build daltonng."""
    p, b1, lam = 3, 0.5, 3.0
    Sp = np.array([[2e-6, 1e-5, 3e-5], [1e-5, 2e-4, 6e-4], [3e-5, 6e-4, 3e-2]])          # a predicted block covariance
    H = np.diag([-b1 * b1 * lam, 0.0, 0.0])                                             # own-block Hessian of the log-pmf
    obs_var = -np.linalg.pinv(H)
    obs_weight = np.where(obs_var != 0, 1.0, 0.0)
    W = np.vstack([np.array([[-0.3, 1.0, 0.0]]), obs_weight])                             # [W~ ; obs_weight]
    V = sl.block_diag(np.zeros((1, 1)), obs_var)
    S = W @ Sp @ W.T + V
    assert np.linalg.matrix_rank(S) == 2 and S.shape == (1 + p, 1 + p)
    assert not S[2:].any() and not S[:, 2:].any()
    lu, piv = sl.lu_factor(S, check_finite=False)
    assert (np.diag(lu)[2:] == 0).all()
    with np.errstate(all="ignore"):
        K = sl.lu_solve((lu, piv), W @ Sp, check_finite=False)
    assert np.isnan(K).all()
