"""Extended-precision (x87 80-bit, 64-bit mantissa) statement of the scalar-measurement dalton recursion.

Used only to measure the float64 *noise floor* of the log-likelihood: the per-step forecast residuals
z = W~ mu_p + d are differences of O(1) quantities that leave ~1e-10 relative rounding noise in z^2/S, so two
faithful float64 evaluations of the same recursion (the reference on CPU vs on GPU, the NumPy oracle vs the CUDA
kernel) legitimately differ by a few 1e-10 of the result.  This module computes the recursion in longdouble
(~2000x finer rounding) so that tests can compare |kernel - exact| with |oracle - exact|.

Restricted to n_bmeas = n_bobs = 1 and interrogate_kramer (the headline config).  Test infrastructure only.
"""
import numpy as np

LD = np.longdouble


def _sym_eig2(a, b, c):
    """eigenvalues (w1, w2) and first eigenvector (cs, sn) of [[a,b],[b,c]], vectorised"""
    tr, df = a + c, a - c
    rt = np.sqrt(df * df + 4 * b * b)
    w1 = (tr + rt) / 2
    w2 = (a * c - b * b) / np.where(w1 == 0, LD(1), w1)
    # eigenvector of w1: (b, w1 - a) or (w1 - c, b), pick the better conditioned
    v0a, v1a = b, w1 - a
    v0b, v1b = w1 - c, b
    use_b = np.abs(v0b) + np.abs(v1b) > np.abs(v0a) + np.abs(v1a)
    v0 = np.where(use_b, v0b, v0a); v1 = np.where(use_b, v1b, v1a)
    nrm = np.sqrt(v0 * v0 + v1 * v1)
    nrm = np.where(nrm == 0, LD(1), nrm)
    return w1, w2, v0 / nrm, v1 / nrm


def dalton_ld(fun, jac, W, X0, t_min, t_max, n_steps, Q, R, theta, obs_data, obs_ind, obs_weight, obs_var):
    """log p(Y|Z) per theta in longdouble; fun/jac take and return longdouble arrays (B, nb, 1[, p])."""
    X0 = X0.astype(LD); theta = theta.astype(LD); Q = Q.astype(LD); R = R.astype(LD); W = W.astype(LD)
    obs_data = obs_data.astype(LD); obs_weight = obs_weight.astype(LD); obs_var = obs_var.astype(LD)
    B, nb, p = X0.shape
    n_obs = len(obs_ind)
    L2PI = np.log(2 * LD(np.pi)) if False else LD("1.8378770664093454835606594728112353")
    cut = LD(1e-8)

    def term(w, z):
        keep = np.abs(w) > cut
        ws = np.where(keep, w, LD(1))
        return np.where(keep, -(z * z / ws + np.log(ws)) / 2 - L2PI / 2, LD(0))

    def step(mu, S, t, obs_i):
        # predict
        mu = np.einsum("nij,bnj->bni", Q, mu)
        S = np.einsum("nij,bnjk,nlk->bnil", Q, S, Q) + R
        f = fun(mu, t, theta); J = jac(mu, t, theta)                      # (B,nb,1), (B,nb,1,p)
        wm = W[None, :, 0, :] - J[:, :, 0, :]                            # (B,nb,p)
        d = -f[:, :, 0] + np.einsum("bnp,bnp->bn", J[:, :, 0, :], mu)
        v1 = np.einsum("bnij,bnj->bni", S, wm)
        s11 = np.einsum("bni,bni->bn", wm, v1)
        r1 = -(np.einsum("bni,bni->bn", wm, mu) + d)
        if obs_i is None:
            lp = term(s11, r1).sum(axis=1)
            K = v1 / s11[..., None]
            mu = mu + K * r1[..., None]
            S = S - K[..., :, None] * v1[..., None, :]
            return mu, S, lp
        D = obs_weight[obs_i][None, :, 0, :]                              # (1,nb,p)
        v2 = np.einsum("bnij,bnj->bni", S, np.broadcast_to(D, mu.shape))
        s12 = np.einsum("bni,bni->bn", wm, v2)
        s22 = np.einsum("bni,bni->bn", np.broadcast_to(D, mu.shape), v2) + obs_var[obs_i][None, :, 0, 0]
        r2 = obs_data[obs_i][None, :, 0] - np.einsum("bni,bni->bn", np.broadcast_to(D, mu.shape), mu)
        w1, w2, cs, sn = _sym_eig2(s11, s12, s22)
        z1 = cs * r1 + sn * r2; z2 = -sn * r1 + cs * r2
        lp = (term(w1, z1) + term(w2, z2)).sum(axis=1)
        det = s11 * s22 - s12 * s12
        k1 = (v1 * s22[..., None] - v2 * s12[..., None]) / det[..., None]
        k2 = (v2 * s11[..., None] - v1 * s12[..., None]) / det[..., None]
        mu = mu + k1 * r1[..., None] + k2 * r2[..., None]
        S = S - k1[..., :, None] * v1[..., None, :] - k2[..., :, None] * v2[..., None, :]
        return mu, S, lp

    ll_zy = np.zeros(B, dtype=LD); ll_z = np.zeros(B, dtype=LD)
    i = 0
    if obs_ind[0] == 0:
        for b in range(nb):
            mz = X0[:, b] @ obs_weight[0, b, 0]
            ll_zy += term(np.broadcast_to(obs_var[0, b, 0, 0], (B,)), obs_data[0, b, 0] - mz)
        i = 1
    mzy, Szy = X0.copy(), np.zeros((B, nb, p, p), dtype=LD)
    mz_, Sz = X0.copy(), np.zeros((B, nb, p, p), dtype=LD)
    for n in range(n_steps):
        t = LD(t_min) + (LD(t_max) - LD(t_min)) * (n + 1) / n_steps
        ic = min(i, n_obs - 1)
        if n + 1 == obs_ind[ic]:
            mzy, Szy, lp = step(mzy, Szy, t, ic); i += 1
        else:
            mzy, Szy, lp = step(mzy, Szy, t, None)
        ll_zy += lp
        mz_, Sz, lp = step(mz_, Sz, t, None)
        ll_z += lp
    return ll_zy - ll_z


def fitz_fun_ld(X, t, th):
    a, b, c = th[:, 0], th[:, 1], th[:, 2]
    V, R = X[:, 0, 0], X[:, 1, 0]
    out = np.empty((X.shape[0], 2, 1), dtype=LD)
    out[:, 0, 0] = c * (V - V * V * V / 3 + R)
    out[:, 1, 0] = -1 / c * (V - a + b * R)
    return out


def fitz_jac_ld(X, t, th):
    a, b, c = th[:, 0], th[:, 1], th[:, 2]
    V = X[:, 0, 0]
    J = np.zeros((X.shape[0], 2, 1, X.shape[2]), dtype=LD)
    J[:, 0, 0, 0] = c * (1 - V * V)
    J[:, 1, 0, 0] = -1 / c * b
    return J
