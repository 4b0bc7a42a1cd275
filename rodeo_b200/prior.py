"""IBM prior (host side) -- same closed forms as reference src/rodeo/prior/ibm.py:21-88.

Q, R are *inputs* to the kernels (tiny, shared by the theta batch); they are built on the host with the
reference's own formula, including ``exp(gammaln(x+1))`` for the factorials.
"""
import math

import numpy as np


def _factorial(x):
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    for idx in np.ndindex(x.shape):
        a = float(x[idx]) + 1.0
        out[idx] = math.inf if (a <= 0.0 and a == math.floor(a)) else math.exp(math.lgamma(a))
    return out


def ibm_state(dt, q, sigma):
    """reference src/rodeo/prior/ibm.py:37-62"""
    I, J = np.meshgrid(np.arange(q + 1), np.arange(q + 1), indexing="ij", sparse=True)
    mesh = (J - I).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        Q = np.nan_to_num(np.float64(dt) ** mesh / _factorial(mesh), nan=0.0)
    mesh = (2.0 * q + 1.0) - I - J
    R = sigma ** 2 * (np.float64(dt) ** mesh) / (mesh * _factorial(q - I) * _factorial(q - J))
    return Q, R


def ibm_init(dt, n_deriv, sigma):
    """(wgt_state, var_state) of shapes (n_block, p, p) -- reference src/rodeo/prior/ibm.py:65-88.

    ``sigma`` of shape (B, n_block) (a theta-dependent prior scale, as in the reference's
    docs/examples/parameter.md:218-222 under vmap) gives var_state of shape (B, n_block, p, p).
    """
    if hasattr(sigma, "detach"):
        sigma = sigma.detach().cpu().numpy()
    sigma = np.asarray(sigma, dtype=np.float64)
    Q1, R1 = ibm_state(dt, n_deriv - 1, 1)
    if sigma.ndim == 2:
        Q = np.repeat(Q1[None], sigma.shape[1], axis=0)
        return Q, sigma[:, :, None, None] ** 2 * R1[None, None]
    Q = np.repeat(Q1[None], len(sigma), axis=0)
    R = np.stack([sigma[b] ** 2 * R1 for b in range(len(sigma))])
    return Q, R


def indep_init(prior_pars):
    """Combine blocks of prior parameters into dense (1, n_block*p, n_block*p) matrices -- reference
    src/rodeo/prior/indep_init.py:8-23.  Host-side helper for code that wants the non-blocked form (the reference's
    examples/solve_nb.py); the kernels themselves run the blocked form and reject a model whose block shape differs
    from the one it was compiled for."""
    import scipy.linalg
    prior_weight, prior_var = (np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float64)
                               for a in prior_pars)
    return scipy.linalg.block_diag(*prior_weight)[None, :], scipy.linalg.block_diag(*prior_var)[None, :]
