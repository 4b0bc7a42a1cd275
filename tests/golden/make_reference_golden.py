#!/usr/bin/env python
"""Regenerate tests/golden/reference_vectors.npz by RUNNING THE REFERENCE'S OWN SOURCE.

The unmodified files under /root/reference/src/rodeo (mlysy/rodeo v1.1.3) are imported and executed with
``oracle/jaxshim`` -- a NumPy stand-in for the handful of jax entry points they use -- first on sys.path (JAX itself
is not installed in this image and there is no network).  Every output below therefore comes from the reference's
own solve.py / interrogate.py / kalmantv/*.py / inference/*.py / prior/ibm.py / utils.py control flow and formulas,
evaluated in float64 through the same LAPACK routines jaxlib's CPU backend calls.  See the shim's docstring for what
this does and does not pin (not pinned: XLA's operation order, JAX's threefry streams).

This script can only run where /root/reference exists (the build container); the vectors it writes are committed and
are what tests/test_reference_golden.py checks the oracle (CPU) and the CUDA path (GPU box) against.

    python tests/golden/make_reference_golden.py
"""
import functools
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RODEO_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))
import problems as P  # noqa: E402


def _import_reference():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "jaxshim"), os.path.join(REF, "src")]
    warnings.simplefilter("ignore", SyntaxWarning)
    import jax
    import jax.numpy as jnp
    import rodeo
    import importlib
    # `import rodeo.inference.dalton as m` would bind the FUNCTION that rodeo/inference/__init__.py re-exports
    rdalton = importlib.import_module("rodeo.inference.dalton")
    rfenrir = importlib.import_module("rodeo.inference.fenrir")
    import rodeo.kalmantv.standard as rstd
    import rodeo.kalmantv.square_root as rsqrt
    return jax, jnp, rodeo, rdalton, rfenrir, rstd, rsqrt


def build():
    jax, jnp, rodeo, rdalton, rfenrir, rstd, rsqrt = _import_reference()
    A = np.asarray
    out = {}

    # ---- ODE right-hand sides exactly as the reference's README / docs write them ---------------------------------
    def fitz_fun(X, t, **params):                      # README.md:92-99
        a, b, c = params["theta"]
        V, R = X[:, 0]
        return jnp.array([[c * (V - V * V * V / 3 + R)],
                          [-1 / c * (V - a + b * R)]])

    def lorenz(X_t, t, theta):                         # docs/examples/lorenz.md:95-101
        rho, sigma, beta = theta
        x, y, z = X_t[:, 0]
        dx = -sigma * x + sigma * y
        dy = rho * x - y - x * z
        dz = -beta * z + x * y
        return jnp.array([[dx], [dy], [dz]])

    def higher_fun(x, t, **params):                    # docs/examples/higher_order.md:47-58, with theta = (omega, k)
        w, k = params["theta"]
        return jnp.array([[jnp.sin(w * t) - k * x[0, 0]]])

    kramer = rodeo.interrogate.interrogate_kramer
    key = jax.random.PRNGKey(0)

    def per_theta(fn, pr, **extra):
        """call an un-batched reference function once per theta and stack (the reference leaves the theta batch to
        the user's jit(vmap))"""
        res = []
        for i in range(pr["theta"].shape[0]):
            r = fn(jnp.array(pr["X0"][i]), jnp.array(pr["theta"][i]), **extra)
            res.append(r)
        if isinstance(res[0], tuple):
            return tuple(np.stack([A(r[k]) for r in res]) for k in range(len(res[0])))
        return np.stack([A(r) for r in res])

    def save_problem(tag, pr, ob=None):
        out[f"{tag}_in_grid"] = np.array([pr["t_min"], pr["t_max"], pr["n_steps"]])
        for k in ("W", "X0", "theta", "Q", "R"):
            out[f"{tag}_in_{k}"] = pr[k]
        if ob is not None:
            for k, v in ob.items():
                out[f"{tag}_in_{k}"] = v

    # ===============================================================================================================
    # FitzHugh-Nagumo, 3 thetas, N = 60 on [0, 3], observations at t = 0, 1, 2, 3
    # ===============================================================================================================
    pr = P.fitz_problem(3, n_steps=60, t_max=3.0, seed=123)
    ob = P.fitz_obs(pr, None, n_obs=4)
    save_problem("fitz", pr, ob)
    # the reference's own helpers must reproduce the inputs: first_order_pad / ibm_init (utils.py:80-102, ibm.py:65-88)
    W, pad = rodeo.utils.first_order_pad(fitz_fun, 2, 3)
    out["fitz_ref_W"] = A(W)
    out["fitz_ref_X0"] = np.stack([A(pad(jnp.array(pr["x0"][i]), 0.0, theta=jnp.array(pr["theta"][i])))
                                   for i in range(3)])
    Qr, Rr = rodeo.prior.ibm_init(dt=3.0 / 60, n_deriv=3, sigma=jnp.array([0.1, 0.1]))
    out["fitz_ref_Q"], out["fitz_ref_R"] = A(Qr), A(Rr)
    pp = (jnp.array(pr["Q"]), jnp.array(pr["R"]))
    common = dict(ode_fun=fitz_fun, ode_weight=jnp.array(pr["W"]), t_min=0.0, t_max=3.0, n_steps=60, prior_pars=pp)
    obs = dict(obs_data=jnp.array(ob["obs_data"]), obs_times=jnp.array(ob["obs_times"]),
               obs_weight=jnp.array(ob["obs_weight"]), obs_var=jnp.array(ob["obs_var"]))

    for name in ("kramer", "schober", "rodeo"):
        interr = getattr(rodeo.interrogate, "interrogate_" + name)
        m, v = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=interr, theta=th, **common), pr)
        out[f"fitz_{name}_mean"], out[f"fitz_{name}_var"] = m, v

    # interrogate_chkrebtii: the shim logs the standard normals of every jax.random.multivariate_normal call, in call
    # order (step-major, block-minor: interrogate.py:23-34 under solve.py's scan)
    chk = functools.partial(rodeo.interrogate.interrogate_chkrebtii, kalman_type="standard")
    zs, ms, vs = [], [], []
    for i in range(3):
        del jax.random.DRAW_LOG[:]
        m, v = rodeo.solve_mv(key=jax.random.PRNGKey(10 + i), ode_init=jnp.array(pr["X0"][i]), interrogate=chk,
                              theta=jnp.array(pr["theta"][i]), **common)
        z = np.stack([e[2] for e in jax.random.DRAW_LOG])
        assert z.shape == (60 * 2, 3) and all(e[1] == "mvn-cholesky" for e in jax.random.DRAW_LOG)
        zs.append(z.reshape(60, 2, 3)); ms.append(A(m)); vs.append(A(v))
    out["fitz_chkrebtii_z"], out["fitz_chkrebtii_mean"], out["fitz_chkrebtii_var"] = np.stack(zs), np.stack(ms), np.stack(vs)

    # solve_sim (kramer forward pass => the only draws are the smoother's, method='svd': solve.py:179-186)
    zs, xs = [], []
    for i in range(3):
        del jax.random.DRAW_LOG[:]
        x = rodeo.solve_sim(key=jax.random.PRNGKey(20 + i), ode_init=jnp.array(pr["X0"][i]), interrogate=kramer,
                            theta=jnp.array(pr["theta"][i]), **common)
        log = jax.random.DRAW_LOG
        assert len(log) == 60 and all(e[1] == "mvn-svd" and e[2].shape == (2, 3) for e in log)
        z = np.zeros((61, 2, 3))
        z[60] = log[0][2]                               # terminal draw first, then t = N-1 .. 1 (reverse scan)
        for k in range(1, 60):
            z[60 - k] = log[k][2]
        zs.append(z); xs.append(A(x))
    out["fitz_sim_z"], out["fitz_sim_x"] = np.stack(zs), np.stack(xs)

    out["fitz_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs), pr)
    out["fitz_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs), pr)

    def obs_loglik(obs_data, ode_data, **params):      # Gaussian measurement model of docs/examples/parameter.md:333-340
        return jnp.sum(jax.scipy.stats.norm.logpdf(obs_data[:, :, 0], ode_data[:, :, 0], np.sqrt(0.005)))
    ll, Xt = per_theta(lambda X0, th: rodeo.inference.basic(
        key=key, ode_init=X0, interrogate=kramer, theta=th, obs_data=obs["obs_data"], obs_times=obs["obs_times"],
        obs_loglik=obs_loglik, **common), pr)
    out["fitz_basic"], out["fitz_basic_Xt"] = ll, Xt

    m, v = per_theta(lambda X0, th: rdalton.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs), pr)
    out["fitz_dalton_mean"], out["fitz_dalton_var"] = m, v
    m, v = per_theta(lambda X0, th: rfenrir.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs), pr)
    out["fitz_fenrir_mean"], out["fitz_fenrir_var"] = m, v

    # observations that do NOT start on t_min and do not end on t_max (dalton.py:207-215 `_no_logy0`, fenrir's
    # terminal `_no_obs` branch, fenrir.py:196-220), and fall between grid points (left insertion of searchsorted)
    ob2 = {k: v[1:3] for k, v in ob.items()}
    ob2["obs_times"] = np.array([0.97, 2.04])
    save_problem("fitzmid", pr, ob2)
    obs2 = {k: jnp.array(v) for k, v in ob2.items()}
    out["fitzmid_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs2), pr)
    out["fitzmid_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs2), pr)

    # two observation rows per block with correlated noise (n_bobs = 2): the stacked 3-row update and the 3x3 / 2x2
    # eigen-decompositions of the forecast variances (dalton.py:136-149, fenrir.py:160-179)
    rngb = np.random.default_rng(77)
    n_ob = len(ob["obs_times"])
    ob3 = {"obs_times": ob["obs_times"].copy()}
    D3 = np.zeros((n_ob, 2, 2, 3)); D3[..., 0, 0] = 1.0; D3[..., 1, 0] = 0.5; D3[..., 1, 1] = 0.2
    Om3 = np.zeros((n_ob, 2, 2, 2)); Om3[..., 0, 0] = 0.005; Om3[..., 1, 1] = 0.02; Om3[..., 0, 1] = Om3[..., 1, 0] = 0.003
    ob3["obs_weight"], ob3["obs_var"] = D3, Om3
    ob3["obs_data"] = np.concatenate([ob["obs_data"], 0.5 * ob["obs_data"] + 0.05 * rngb.standard_normal(ob["obs_data"].shape)],
                                     axis=2)
    save_problem("fitzbobs2", pr, ob3)
    obs3 = {k: jnp.array(v) for k, v in ob3.items()}
    out["fitzbobs2_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs3), pr)
    out["fitzbobs2_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obs3), pr)

    # shortest solves the reference's scans support (the backward scan has length n_steps - 1 >= 1): n_steps = 2, 3,
    # with observations on both end points
    for Ns in (2, 3):
        prs = P.fitz_problem(3, n_steps=Ns, t_max={2: 0.2, 3: 0.3}[Ns], seed=5)
        obx = P.fitz_obs(prs, None, n_obs=2)
        save_problem(f"fitzN{Ns}", prs, obx)
        cs = dict(ode_fun=fitz_fun, ode_weight=jnp.array(prs["W"]), t_min=0.0, t_max=prs["t_max"], n_steps=Ns,
                  prior_pars=(jnp.array(prs["Q"]), jnp.array(prs["R"])))
        obj = {k: jnp.array(v) for k, v in obx.items()}
        m, v = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **cs), prs)
        out[f"fitzN{Ns}_mean"], out[f"fitzN{Ns}_var"] = m, v
        out[f"fitzN{Ns}_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
            key=key, ode_init=X0, interrogate=kramer, theta=th, **cs, **obj), prs)
        out[f"fitzN{Ns}_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
            key=key, ode_init=X0, interrogate=kramer, theta=th, **cs, **obj), prs)

    # an observation time just past t_max: searchsorted gives n_steps + 1, an index no step ever equals
    obp = {k: v[1:3].copy() for k, v in ob.items()}
    obp["obs_times"] = np.array([1.0, 3.0 + 1e-9])
    save_problem("fitzpast", pr, obp)
    obsp = {k: jnp.array(v) for k, v in obp.items()}
    out["fitzpast_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obsp), pr)
    out["fitzpast_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common, **obsp), pr)

    # square-root Kalman family (solve.py:236-241 with kalman_type="square-root"; docs/examples/higher_order.md:108-112)
    chol = jax.vmap(jnp.linalg.cholesky)(pp[1])
    common_sq = dict(common, prior_pars=(pp[0], chol))
    m, L = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th,
                                                   kalman_type="square-root", **common_sq), pr)
    out["fitz_sqrt_mean"], out["fitz_sqrt_var"] = m, L @ np.swapaxes(L, -1, -2)      # compare L L^T (QR sign freedom)

    # ===============================================================================================================
    # BASELINE configs[0]: the README walkthrough itself (single theta, N = 800 on [0, 40])
    # ===============================================================================================================
    pr1 = P.fitz_problem(1, jitter=False)
    ob1 = P.fitz_obs(pr1, None, n_obs=41)
    save_problem("readme", pr1, ob1)
    common1 = dict(ode_fun=fitz_fun, ode_weight=jnp.array(pr1["W"]), t_min=0.0, t_max=40.0, n_steps=800,
                   prior_pars=(jnp.array(pr1["Q"]), jnp.array(pr1["R"])))
    obs1 = {k: jnp.array(v) for k, v in ob1.items()}
    m, v = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **common1), pr1)
    out["readme_mean"], out["readme_var"] = m, v
    out["readme_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common1, **obs1), pr1)
    out["readme_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common1, **obs1), pr1)

    # ===============================================================================================================
    # Lorenz63 (docs/examples/lorenz.md), 2 thetas, N = 100 on [0, 0.5]; sigma = 5e7 as in the docs
    # ===============================================================================================================
    prl = P.lorenz_problem(2, n_steps=100, t_max=0.5, seed=7)
    save_problem("lorenz", prl)
    commonl = dict(ode_fun=lorenz, ode_weight=jnp.array(prl["W"]), t_min=0.0, t_max=0.5, n_steps=100,
                   prior_pars=(jnp.array(prl["Q"]), jnp.array(prl["R"])))
    m, v = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **commonl), prl)
    out["lorenz_mean"], out["lorenz_var"] = m, v

    # ===============================================================================================================
    # second-order ODE x'' = sin(w t) - k x  (docs/examples/higher_order.md), n_block = 1, n_bstate = 4
    # ===============================================================================================================
    pr2 = P.second_order_problem(2, n_steps=80, t_max=4.0, sigma=0.1, seed=123)
    ob2 = P.second_order_obs(pr2, n_obs=5)
    save_problem("so", pr2, ob2)
    common2 = dict(ode_fun=higher_fun, ode_weight=jnp.array(pr2["W"]), t_min=0.0, t_max=4.0, n_steps=80,
                   prior_pars=(jnp.array(pr2["Q"]), jnp.array(pr2["R"])))
    obs2 = {k: jnp.array(v) for k, v in ob2.items()}
    m, v = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **common2), pr2)
    out["so_mean"], out["so_var"] = m, v
    out["so_fenrir"] = per_theta(lambda X0, th: rodeo.inference.fenrir(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common2, **obs2), pr2)
    out["so_dalton"] = per_theta(lambda X0, th: rodeo.inference.dalton(
        key=key, ode_init=X0, interrogate=kramer, theta=th, **common2, **obs2), pr2)

    # ===============================================================================================================
    # Hes1 (log scale) and SEIRAH, right-hand sides and settings of the reference's examples/timings.py:251-258, 339-352
    # ===============================================================================================================
    def hes1(X, t, theta):
        P_, M_, H_ = jnp.exp(X[:, 0])
        a, b, c, d, e, f, g = theta
        logP = -a * H_ + b * M_ / P_ - c
        logM = -d + e / (1 + P_ * P_) / M_
        logH = -a * P_ + f / (1 + P_ * P_) / H_ - g
        return jnp.array([[logP], [logM], [logH]])

    def seirah(X, t, theta):
        S, E, I, R, A_, H = X[:, 0]
        N = S + E + I + R + A_ + H
        b, r, alpha, D_e, D_I, D_q = theta
        D_h = 30
        dS = -b * S * (I + alpha * A_) / N
        dE = b * S * (I + alpha * A_) / N - E / D_e
        dI = r * E / D_e - I / D_q - I / D_I
        dR = (I + A_) / D_I + H / D_h
        dA = (1 - r) * E / D_e - A_ / D_I
        dH = I / D_q - H / D_h
        return jnp.array([[dS], [dE], [dI], [dR], [dA], [dH]])

    for tag, fun, nv, x0v, th0, t_max_, n_st, seed in (
            ("hes1", hes1, 3, np.log([1.439, 2.037, 17.904]), [0.022, 0.3, 0.031, 0.028, 0.5, 20, 0.3], 240.0, 120, 5),
            ("seirah", seirah, 6, [63804435., 15492., 21752., 0., 618013., 93583.], [2.23, 0.034, 0.55, 5.1, 2.3, 1.13],
             60.0, 80, 6)):
        rng_ = np.random.default_rng(seed)
        theta_ = np.asarray(th0) * np.exp(0.02 * rng_.standard_normal((2, len(th0))))
        Wm, padm = rodeo.utils.first_order_pad(fun, nv, 3)
        X0_ = np.stack([A(padm(jnp.array(np.asarray(x0v, dtype=float)), 0.0, theta=jnp.array(theta_[i]))) for i in range(2)])
        Qm, Rm = rodeo.prior.ibm_init(dt=t_max_ / n_st, n_deriv=3, sigma=jnp.array([0.1] * nv))
        prm = dict(W=A(Wm), X0=X0_, theta=theta_, Q=A(Qm), R=A(Rm), t_min=0.0, t_max=t_max_, n_steps=n_st)
        save_problem(tag, prm)
        cm = dict(ode_fun=fun, ode_weight=Wm, t_min=0.0, t_max=t_max_, n_steps=n_st, prior_pars=(Qm, Rm))
        m, v = per_theta(lambda X0, th: rodeo.solve_mv(key=key, ode_init=X0, interrogate=kramer, theta=th, **cm), prm)
        out[f"{tag}_mean"], out[f"{tag}_var"] = m, v

    # ===============================================================================================================
    # sigma as part of theta: the prior is rebuilt per theta with the reference's ibm_init (docs/examples/parameter.md:
    # 218-222), here on FitzHugh-Nagumo with the dalton log-likelihood and solve_mv
    # ===============================================================================================================
    sig = 0.1 * np.exp(0.3 * np.random.default_rng(3).standard_normal((3, 2)))
    out["fitzsig_in_sigma"] = sig
    ms_, vs_, ll_ = [], [], []
    for i in range(3):
        Qi, Ri = rodeo.prior.ibm_init(dt=3.0 / 60, n_deriv=3, sigma=jnp.array(sig[i]))
        ci = dict(common, prior_pars=(Qi, Ri))
        m, v = rodeo.solve_mv(key=key, ode_init=jnp.array(pr["X0"][i]), interrogate=kramer, theta=jnp.array(pr["theta"][i]), **ci)
        ll_.append(float(rodeo.inference.dalton(key=key, ode_init=jnp.array(pr["X0"][i]), interrogate=kramer,
                                                theta=jnp.array(pr["theta"][i]), **ci, **obs)))
        ms_.append(A(m)); vs_.append(A(v))
    out["fitzsig_mean"], out["fitzsig_var"], out["fitzsig_dalton"] = np.stack(ms_), np.stack(vs_), np.array(ll_)

    # ===============================================================================================================
    # a general per-theta prior: every theta brings its own (Q, R), neither shared nor a multiple of a shared matrix
    # (the reference takes whatever prior_pars the user's vmap hands it, docs/examples/parameter.md:218-236)
    # ===============================================================================================================
    rq = np.random.default_rng(31)
    Qg = np.repeat(pr["Q"][None], 3, 0).copy()
    Rg = np.repeat(pr["R"][None], 3, 0).copy()
    for i in range(3):
        for b in range(2):
            Qg[i, b] = Qg[i, b] * (1.0 + 0.05 * rq.standard_normal((3, 3))) + 0.01 * rq.standard_normal((3, 3))
            Lr = np.linalg.cholesky(Rg[i, b])
            a_ = 0.3 * rq.standard_normal((3, 3))
            Rg[i, b] = Lr @ (np.eye(3) + a_ @ a_.T) @ Lr.T
            Rg[i, b] = 0.5 * (Rg[i, b] + Rg[i, b].T)
    out["fitzqr_in_Q"], out["fitzqr_in_R"] = Qg, Rg
    ms_, vs_, ll_, fl_ = [], [], [], []
    for i in range(3):
        ci = dict(common, prior_pars=(jnp.array(Qg[i]), jnp.array(Rg[i])))
        a_ = dict(key=key, ode_init=jnp.array(pr["X0"][i]), interrogate=kramer, theta=jnp.array(pr["theta"][i]))
        m, v = rodeo.solve_mv(**a_, **ci)
        ll_.append(float(rodeo.inference.dalton(**a_, **ci, **obs)))
        fl_.append(float(rodeo.inference.fenrir(**a_, **ci, **obs)))
        ms_.append(A(m)); vs_.append(A(v))
    out["fitzqr_mean"], out["fitzqr_var"] = np.stack(ms_), np.stack(vs_)
    out["fitzqr_dalton"], out["fitzqr_fenrir"] = np.array(ll_), np.array(fl_)

    # ===============================================================================================================
    # n_bmeas = 2: ONE block that holds two variables, state (x, x', x'', y, y', y''), measurement rows
    # W = [[0,1,0,0,0,0],[0,0,0,0,1,0]], prior = blockdiag of two IBM priors; x' = -a x + sin t, y' = -b y^2
    # (theta = (a, b, c)); observations of x only (n_bobs = 1).  The two variables are not coupled: with a coupled pair
    # (FitzHugh-Nagumo in one block) the reference's covariance-form recursion itself loses the symmetry of its variance
    # within the 60 steps and its dalton value is NaN, so there is nothing to pin.
    # ===============================================================================================================
    def pair_one_block(X, t, **params):
        a, b, c = params["theta"]
        return jnp.array([[-a * X[0, 0] + jnp.sin(t), -b * X[0, 3] * X[0, 3]]])

    Q3, R3 = rodeo.prior.ibm_init(dt=3.0 / 60, n_deriv=3, sigma=jnp.array([0.1, 0.1]))
    Q1b = np.zeros((1, 6, 6)); R1b = np.zeros((1, 6, 6))
    Q1b[0, :3, :3], Q1b[0, 3:, 3:] = A(Q3)[0], A(Q3)[1]
    R1b[0, :3, :3], R1b[0, 3:, 3:] = A(R3)[0], A(R3)[1]
    W1b = np.zeros((1, 2, 6)); W1b[0, 0, 1] = 1.0; W1b[0, 1, 4] = 1.0
    X01b = np.zeros((3, 1, 6))
    X01b[:, 0, 0], X01b[:, 0, 3] = 1.0 + 0.1 * np.arange(3), 0.5 - 0.05 * np.arange(3)
    X01b[:, 0, 1] = -pr["theta"][:, 0] * X01b[:, 0, 0]
    X01b[:, 0, 4] = -pr["theta"][:, 1] * X01b[:, 0, 3] ** 2
    ow1b = np.zeros((4, 1, 1, 6)); ow1b[..., 0] = 1.0
    ob1b = dict(obs_data=ob["obs_data"][:, 0:1, :], obs_times=ob["obs_times"], obs_weight=ow1b,
                obs_var=ob["obs_var"][:, 0:1])
    out["pair1b_in_W"], out["pair1b_in_Q"], out["pair1b_in_R"], out["pair1b_in_X0"] = W1b, Q1b, R1b, X01b
    for k, v in ob1b.items():
        out["pair1b_in_" + k] = v
    c1b = dict(ode_fun=pair_one_block, ode_weight=jnp.array(W1b), t_min=0.0, t_max=3.0, n_steps=60,
               prior_pars=(jnp.array(Q1b), jnp.array(R1b)))
    o1b = {k: jnp.array(v) for k, v in ob1b.items()}
    ms_, vs_, ll_, fl_ = [], [], [], []
    for i in range(3):
        a_ = dict(key=key, ode_init=jnp.array(X01b[i]), interrogate=kramer, theta=jnp.array(pr["theta"][i]))
        m, v = rodeo.solve_mv(**a_, **c1b)
        ll_.append(float(rodeo.inference.dalton(**a_, **c1b, **o1b)))
        fl_.append(float(rodeo.inference.fenrir(**a_, **c1b, **o1b)))
        ms_.append(A(m)); vs_.append(A(v))
    out["pair1b_mean"], out["pair1b_var"] = np.stack(ms_), np.stack(vs_)
    out["pair1b_dalton"], out["pair1b_fenrir"] = np.array(ll_), np.array(fl_)

    # ===============================================================================================================
    # MAGI log-density (inference/magi.py): trajectories = the kramer posterior means above plus noise, expanded as the
    # ODE prescribes -- FitzHugh-Nagumo (x, f(x, theta), 0), second-order ODE (x, x', sin(w t) - k x, 0).
    # n_active = 1 is the well-conditioned case.  With n_active >= 2 the reference's covariance-form recursion observes
    # two or three entries of a block without noise, loses the symmetry of its variance within ~10 steps and becomes
    # rounding-dominated (two float64 evaluations of the same formulas differ by ~1e-2): those values are stored as
    # *_illcond and only checked loosely.
    # ===============================================================================================================
    rngm = np.random.default_rng(17)
    U = out["fitz_kramer_mean"][:, :, :, 0:1] + 0.01 * rngm.standard_normal((3, 61, 2, 1))
    out["magi_fitz_in_U"] = U

    def fitz_expand(U_, **params):                     # (N+1, nb, 1) -> (N+1, nb, 3)
        f = jax.vmap(lambda u: fitz_fun(u, 0.0, **params)[:, 0])(U_)
        return jnp.concatenate([U_, f[:, :, None], jnp.zeros(U_.shape)], axis=2)

    def magi(U_, expand, na, prior, th):
        return float(rodeo.inference.magi_logdens(jnp.array(U_), expand, na, prior, "standard", theta=jnp.array(th)))
    out["magi_fitz"] = np.array([magi(U[i], fitz_expand, 1, pp, pr["theta"][i]) for i in range(3)])
    out["magi_fitz_illcond"] = np.array([magi(U[i], fitz_expand, 2, pp, pr["theta"][i]) for i in range(3)])
    out["magi_fitz_state"] = np.stack([A(fitz_expand(jnp.array(U[i]), theta=jnp.array(pr["theta"][i]))) for i in range(3)])
    U2 = out["so_mean"][:, :, :, 0:2] + 0.01 * rngm.standard_normal((2, 81, 1, 2))
    out["magi_so_in_U"] = U2

    def so_expand(U_, **params):                       # (x, x') -> (x, x', sin(w t) - k x, 0)
        w, k = params["theta"]
        t = jnp.linspace(0.0, 4.0, 81)
        xdd = jnp.sin(w * t)[:, None] - k * U_[:, :, 0]
        return jnp.concatenate([U_, xdd[:, :, None], jnp.zeros(U_[:, :, 0:1].shape)], axis=2)
    pp2 = (jnp.array(pr2["Q"]), jnp.array(pr2["R"]))
    out["magi_so"] = np.array([magi(U2[i], so_expand, 1, pp2, pr2["theta"][i]) for i in range(2)])
    out["magi_so_illcond"] = np.array([magi(U2[i], so_expand, 3, pp2, pr2["theta"][i]) for i in range(2)])
    out["magi_so_state"] = np.stack([A(so_expand(jnp.array(U2[i]), theta=jnp.array(pr2["theta"][i]))) for i in range(2)])

    # ===============================================================================================================
    # Kalman primitives on random inputs (kalmantv/standard.py), incl. the log-pdf's 1e-8 eigenvalue cut-off
    # ===============================================================================================================
    rng = np.random.default_rng(99)
    n, p, m_ = 6, 4, 2

    def spd(k, d):
        a = rng.standard_normal((k, d, d))
        return a @ np.swapaxes(a, -1, -2) + 0.5 * np.eye(d)
    kin = dict(mu=rng.standard_normal((n, p)), S=spd(n, p), c=rng.standard_normal((n, p)),
               Qm=rng.standard_normal((n, p, p)), Rm=spd(n, p), xm=rng.standard_normal((n, m_)),
               d=rng.standard_normal((n, m_)), Wm=rng.standard_normal((n, m_, p)), Vm=spd(n, m_),
               xn=rng.standard_normal((n, p)), mun=rng.standard_normal((n, p)), Sn=spd(n, p))
    for k, v in kin.items():
        out["kal_in_" + k] = v
    res = {k: [] for k in ("pm", "pS", "um", "uS", "fm", "fS", "sm", "sS", "ssm", "ssS", "cA", "cb", "cV", "lp")}
    for i in range(n):
        g = {k: jnp.array(v[i]) for k, v in kin.items()}
        pm, pS = rstd.predict(g["mu"], g["S"], g["c"], g["Qm"], g["Rm"])
        um, uS = rstd.update(pm, pS, g["xm"], g["d"], g["Wm"], g["Vm"])
        fm, fS = rstd.forecast(pm, pS, g["d"], g["Wm"], g["Vm"])
        sm, sS = rstd.smooth_mv(g["mun"], g["Sn"], um, uS, pm, pS, g["Qm"])
        ssm, ssS = rstd.smooth_sim(g["xn"], um, uS, pm, pS, g["Qm"])
        cA, cb, cV = rstd.smooth_cond(um, uS, pm, pS, g["Qm"])
        lp = rodeo.utils.multivariate_normal_logpdf(g["xm"], fm, fS)
        for k, v in zip(res, (pm, pS, um, uS, fm, fS, sm, sS, ssm, ssS, cA, cb, cV, lp)):
            res[k].append(A(v))
    for k, v in res.items():
        out["kal_" + k] = np.stack(v)
    # eigenvalues around the absolute 1e-8 cut-off (utils.py:74)
    cov = np.stack([np.diag([w, 2.0]) for w in (0.5e-8, 0.99e-8, 1.01e-8, 2e-8, 0.0)])
    x = np.tile(np.array([1e-4, 0.3]), (5, 1))
    out["lpcut_in_cov"], out["lpcut_in_x"] = cov, x
    out["lpcut"] = np.array([float(rodeo.utils.multivariate_normal_logpdf(jnp.array(x[i]), jnp.zeros(2), jnp.array(cov[i])))
                             for i in range(5)])
    return out


if __name__ == "__main__":
    vec = build()
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **vec)
    print("wrote", path, os.path.getsize(path), "bytes,", len(vec), "arrays")
