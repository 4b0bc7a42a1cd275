"""GPU parity: the CUDA path (through the C ABI, via the drop-in Python API) against the NumPy oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): 1e-10 relative in float64, measured as max|a-b| / max|b| per
output array (log-likelihoods: |a-b| / max(1, |b|) per theta).
"""
import functools

import numpy as np
import pytest

import problems as P
from oracle import rodeo_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def rb():
    import rodeo_b200
    from rodeo_b200 import _lib
    _lib.load()
    return rodeo_b200


def _interr(rb, name):
    f = getattr(rb.interrogate, "interrogate_" + name)
    return functools.partial(f, kalman_type="standard") if name == "chkrebtii" else f


ORC_INTERR = {"kramer": orc.interrogate_kramer, "schober": orc.interrogate_schober,
              "rodeo": orc.interrogate_rodeo, "chkrebtii": orc.interrogate_chkrebtii}


def _np(t):
    return t.detach().cpu().numpy()


def ll_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


@pytest.mark.parametrize("interr", ["kramer", "schober", "rodeo"])
def test_solve_mv_fitz(rb, interr):
    pr = P.fitz_problem(64, n_steps=200, t_max=10.0, seed=3)
    m, v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"],
                       _interr(rb, interr), prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])
    om, ov = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"],
                          ORC_INTERR[interr], (pr["Q"], pr["R"]), pr["theta"])
    assert m.shape == om.shape and v.shape == ov.shape
    assert P.maxnorm_rel(_np(m), om) < TOL
    assert P.maxnorm_rel(_np(v), ov) < TOL


def test_solve_mv_readme_config_single_theta(rb):
    # BASELINE configs[0]: README walkthrough, N=800, t in [0,40], single un-batched theta
    pr = P.fitz_problem(1, jitter=False)
    m, v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"][0], 0.0, 40.0, 800,
                       rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"][0])
    om, ov = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 40.0, 800,
                          orc.interrogate_kramer, (pr["Q"], pr["R"]), pr["theta"])
    assert m.shape == (801, 2, 3) and v.shape == (801, 2, 3, 3)
    assert P.maxnorm_rel(_np(m), om[0]) < TOL and P.maxnorm_rel(_np(v), ov[0]) < TOL
    # legacy keyword spelling (reference <= 1.1.2 / BASELINE north_star)
    m2, _ = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"][0], 0.0, 40.0, 800,
                        rb.interrogate.interrogate_kramer, prior_weight=pr["Q"], prior_var=pr["R"],
                        theta=pr["theta"][0])
    assert np.array_equal(_np(m2), _np(m))


def _dalton_pair(rb, pr, ob, interr):
    got = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"],
                              _interr(rb, interr), prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"], **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"],
                      ORC_INTERR[interr], (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"],
                      ob["obs_weight"], ob["obs_var"])
    return _np(got), want


def _fitz_truth_obs(pr, n_obs):
    N, tm = pr["n_steps"], pr["t_max"]
    truth, _ = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], P.fitz_problem(1, N, tm, jitter=False)["X0"],
                            0.0, tm, N, orc.interrogate_kramer, (pr["Q"], pr["R"]), np.array([[0.2, 0.2, 3.0]]))
    return P.fitz_obs(pr, truth[0], n_obs=n_obs)


@pytest.mark.parametrize("N,t_max,n_obs", [(400, 20.0, 21), (800, 40.0, 41)])
def test_dalton_fitz_kramer_against_oracle_and_extended_precision(rb, N, t_max, n_obs):
    """Per-theta gate of tests/noise_floor.py: |kernel - exact| <= 2 |oracle - exact| + 1e-10, `exact` = the C
    restatement in x87 long double, `oracle` = the NumPy oracle (float64, LAPACK)."""
    import noise_floor as NF
    from oracle import c_port
    pr = P.fitz_problem(96, n_steps=N, t_max=t_max, seed=5)
    ob = _fitz_truth_obs(pr, n_obs)
    got, want = _dalton_pair(rb, pr, ob, "kramer")
    ind = orc.obs_index(0.0, t_max, N, ob["obs_times"])
    exact = c_port.dalton_ld("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"], 0.0, t_max, N, pr["Q"], pr["R"],
                             pr["theta"], ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"])
    st = NF.gate(got, want, exact)
    NF.report(f"dalton FN kramer N={N} B=96", st)
    assert got.shape == (96,) and st["ok"], st


def test_dalton_fitz_rodeo_interrogation(rb):
    pr = P.fitz_problem(64, n_steps=400, t_max=20.0, seed=6)
    ob = _fitz_truth_obs(pr, 21)
    got, want = _dalton_pair(rb, pr, ob, "rodeo")
    assert ll_err(got, want) < 2e-9     # same noise floor as above


def test_dalton_eigen_cutoff_regime(rb):
    # tests' set-up sigma=.001, dt=.05: most forecast variances fall below the 1e-8 cut-off and are dropped
    pr = P.fitz_problem(32, n_steps=200, t_max=10.0, sigma=0.001, seed=7)
    ob = P.fitz_obs(pr, None, n_obs=11)
    got = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200,
                              rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                              **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    assert ll_err(_np(got), want) < 1e-9


def test_dalton_ragged_observation_times(rb):
    # first obs after t_min, obs inside grid cells, two obs mapping to the same grid point, last before t_max
    pr = P.fitz_problem(16, n_steps=100, t_max=5.0, seed=9)
    ob = P.fitz_obs(pr, None, n_obs=6)
    ob["obs_times"] = np.array([0.3, 1.0, 1.01, 2.5, 2.5, 4.2])
    got = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, 100,
                              rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                              **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, 100, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    assert ll_err(_np(got), want) < TOL


def test_fenrir_second_order(rb):
    pr = P.second_order_problem(48, n_steps=500, seed=2)
    ob = P.second_order_obs(pr)
    got = rb.inference.fenrir(None, rb.models.second_order_sin, pr["W"], pr["X0"], 0.0, 10.0, 500,
                              rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                              **ob)
    want = orc.fenrir(orc.MODELS["second_order_sin"], pr["W"], pr["X0"], 0.0, 10.0, 500, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    assert ll_err(_np(got), want) < TOL


@pytest.mark.parametrize("N", [1, 2, 7, 16, 17, 33, 100])
def test_fenrir_forcing_buffer_chunk_edges(rb, monkeypatch, N):
    """The warp-specialised fenrir kernel evaluates the right-hand side's forcing term ahead, in 16-step chunks handed
    over with named barriers: step counts below / at / just above a chunk and ragged batches must give bitwise the
    one-warp kernel's log-likelihoods (the same functions in the same order)."""
    import torch
    for B in (5, 48):
        pr = P.second_order_problem(B, n_steps=N, t_max=0.05 * N, sigma=0.1, seed=20 + N)
        obs_t = np.array([0.0, 0.05 * N])
        D = np.zeros((2, 1, 1, 4)); D[..., 0] = 1.0
        ob = dict(obs_data=np.array([[[-1.0]], [[-0.9]]]), obs_times=obs_t, obs_weight=D, obs_var=np.full((2, 1, 1, 1), 0.01))
        run = lambda: rb.inference.fenrir(None, rb.models.second_order_sin, pr["W"], pr["X0"], 0.0, 0.05 * N, N,
                                          rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]),
                                          theta=pr["theta"], **ob)
        monkeypatch.setenv("RODEO_FENRIR_WS", "0")
        one = run()
        monkeypatch.setenv("RODEO_FENRIR_WS", "1")
        ws = run()
        assert bool(torch.isfinite(ws).all()) and torch.equal(one, ws)
        want = orc.fenrir(orc.MODELS["second_order_sin"], pr["W"], pr["X0"], 0.0, 0.05 * N, N, orc.interrogate_kramer,
                          (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                          ob["obs_var"])
        assert ll_err(_np(ws), want) < 1e-9


def test_fenrir_fitz(rb):
    pr = P.fitz_problem(40, n_steps=200, t_max=10.0, seed=11)
    ob = P.fitz_obs(pr, None, n_obs=11)
    got = rb.inference.fenrir(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200,
                              rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                              **ob)
    want = orc.fenrir(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    assert ll_err(_np(got), want) < TOL


def test_solve_sim_injected_normals(rb):
    # deterministic sampling-path parity: same normals, same factor definition on both sides
    pr = P.fitz_problem(32, n_steps=120, t_max=6.0, seed=13)
    rng = np.random.default_rng(0)
    zs = rng.standard_normal((32, 121, 2, 3))
    zi = rng.standard_normal((32, 120, 1, 2, 3))
    for interr in ("kramer", "chkrebtii"):
        x = rb.solve_sim(0, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 6.0, 120, _interr(rb, interr),
                         prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"], _z_smooth=zs, _z_interr=zi)
        kw = dict(factor="ldl") if interr == "chkrebtii" else {}
        # oracle: chkrebtii draws with Cholesky (== ldl factor for SPD), smoothing draws with the ldl factor
        want = orc.solve_sim(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 6.0, 120,
                             functools.partial(ORC_INTERR[interr], **kw) if kw else ORC_INTERR[interr],
                             (pr["Q"], pr["R"]), pr["theta"], z_smooth=zs, z_interrogate=zi[:, :, 0], factor="ldl")
        # the smoothing covariances are singular: entries that are exactly-zero pivots in exact arithmetic are
        # rounding noise, so draws agree to ~1e-8 of the state scale rather than 1e-10
        assert P.maxnorm_rel(_np(x), want) < 1e-7, interr


def test_first_order_pad_on_device(rb):
    pr = P.fitz_problem(8, seed=1)
    W, init = rb.utils.first_order_pad(rb.models.fitzhugh_nagumo, 2, 3)
    X0 = init(pr["x0"], 0.0, theta=pr["theta"])
    assert np.array_equal(W, pr["W"])
    assert P.maxnorm_rel(_np(X0), pr["X0"]) < 1e-15


# ---- other models / code paths -------------------------------------------------------------------------------------
def _mv_pair(rb, name, pr, interr="kramer"):
    m, v = rb.solve_mv(None, getattr(rb.models, name), pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"],
                       _interr(rb, interr), prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])
    om, ov = orc.solve_mv(orc.MODELS[name], pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"], ORC_INTERR[interr],
                          (pr["Q"], pr["R"]), pr["theta"])
    return _np(m), _np(v), om, ov


def test_solve_mv_lorenz63(rb):
    # chaotic: keep the horizon short enough that rounding differences are not amplified past the gate
    pr = P.lorenz_problem(24, n_steps=400, t_max=2.0, seed=4)
    m, v, om, ov = _mv_pair(rb, "lorenz63", pr)
    assert P.maxnorm_rel(m, om) < 1e-9 and P.maxnorm_rel(v, ov) < 1e-9


def test_solve_mv_second_order_p4(rb):
    pr = P.second_order_problem(24, n_steps=400, sigma=0.01, seed=4)
    m, v, om, ov = _mv_pair(rb, "second_order_sin", pr)
    assert P.maxnorm_rel(m, om) < TOL and P.maxnorm_rel(v, ov) < TOL


def _generic_first_order(name, B, n_steps, t_max, x0, theta0, sigma, seed):
    mdl = orc.MODELS[name]
    rng = np.random.default_rng(seed)
    theta = np.asarray(theta0) * np.exp(0.02 * rng.standard_normal((B, len(theta0))))
    W, init = orc.first_order_pad(mdl, mdl.n_block, 3)
    X0 = init(np.tile(np.asarray(x0, dtype=float), (B, 1)), 0.0, theta)
    Q, R = orc.ibm_init(t_max / n_steps, 3, np.array([sigma] * mdl.n_block))
    return dict(model=name, W=W, X0=X0, theta=theta, Q=Q, R=R, t_max=t_max, n_steps=n_steps)


def test_solve_mv_hes1_dual_number_jacobian(rb):
    # reference examples/timings.py:288-300: x0 = log(1.439, 2.037, 17.904), theta below
    pr = _generic_first_order("hes1", 16, 240, 240.0, np.log([1.439, 2.037, 17.904]),
                              [0.022, 0.3, 0.031, 0.028, 0.5, 20, 0.3], 0.1, 5)
    m, v, om, ov = _mv_pair(rb, "hes1", pr)
    assert P.maxnorm_rel(m, om) < TOL and P.maxnorm_rel(v, ov) < TOL


def test_solve_mv_seirah_six_blocks(rb):
    # reference examples/timings.py:405-420
    pr = _generic_first_order("seirah", 16, 80, 60.0, [63804435., 15492., 21752., 0., 618013., 93583.],
                              [2.23, 0.034, 0.55, 5.1, 2.3, 1.13], 0.1, 6)
    m, v, om, ov = _mv_pair(rb, "seirah", pr)
    assert P.maxnorm_rel(m, om) < TOL and P.maxnorm_rel(v, ov) < 1e-9


def test_dense_prior_and_general_weight_path(rb):
    # a non-IBM prior (dense Q) and a W that is not a unit row force the general instantiation
    pr = P.fitz_problem(32, n_steps=150, t_max=7.5, seed=8)
    rng = np.random.default_rng(1)
    Q = pr["Q"] + 0.01 * rng.standard_normal(pr["Q"].shape)
    W = pr["W"].copy(); W[:, 0, 2] = 0.05; W[:, 0, 0] = -0.02
    pr2 = dict(pr, Q=Q, W=W)
    m, v, om, ov = _mv_pair(rb, "fitzhugh_nagumo", pr2)
    assert P.maxnorm_rel(m, om) < TOL and P.maxnorm_rel(v, ov) < TOL
    ob = P.fitz_obs(pr2, None, n_obs=6)
    got = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, W, pr["X0"], 0.0, 7.5, 150,
                              rb.interrogate.interrogate_kramer, prior_pars=(Q, pr["R"]), theta=pr["theta"], **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], W, pr["X0"], 0.0, 7.5, 150, orc.interrogate_kramer,
                      (Q, pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert ll_err(_np(got), want) < 2e-9


def test_fenrir_observations_on_both_ends(rb):
    # obs at t_min (scan reaches t = 0) and at t_max (terminal-point update, fenrir.py:196-220)
    pr = P.fitz_problem(24, n_steps=100, t_max=5.0, seed=12)
    ob = P.fitz_obs(pr, None, n_obs=6)
    assert ob["obs_times"][0] == 0.0 and ob["obs_times"][-1] == 5.0
    got = rb.inference.fenrir(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, 100,
                              rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                              **ob)
    want = orc.fenrir(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, 100, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    assert ll_err(_np(got), want) < TOL
    # interior-only observations
    ob["obs_times"] = np.array([0.7, 1.3, 2.0, 2.9, 3.3, 4.1])
    got = rb.inference.fenrir(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, 100,
                              rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                              **ob)
    want = orc.fenrir(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, 100, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    assert ll_err(_np(got), want) < TOL


def test_basic_returns_loglik_and_trajectory(rb):
    import torch
    pr = P.fitz_problem(20, n_steps=100, t_max=5.0, seed=14)
    ob = P.fitz_obs(pr, None, n_obs=6)
    Y = ob["obs_data"][:, :, 0]

    def loglik_t(obs_data, ode_data, **params):      # user function on torch tensors (reference README.md:186-200)
        y = torch.as_tensor(Y, device=ode_data.device)
        return (-0.5 * ((y - ode_data[..., 0]) / 0.07) ** 2).sum(dim=(-1, -2))

    ll, Xt = rb.inference.basic(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, 100,
                                rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]),
                                obs_data=ob["obs_data"], obs_times=ob["obs_times"], obs_loglik=loglik_t,
                                theta=pr["theta"])
    oll, oXt = orc.basic(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, 100, orc.interrogate_kramer,
                         (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"],
                         lambda od, xd, th: (-0.5 * ((Y - xd[..., 0]) / 0.07) ** 2).sum(axis=(-1, -2)))
    assert P.maxnorm_rel(_np(Xt), oXt) < TOL and P.maxnorm_rel(_np(ll), oll) < 1e-9


# ---- sampling paths: distribution-level and determinism ------------------------------------------------------------------
def test_solve_sim_draws_have_solve_mv_moments_for_a_linear_ode(rb):
    """For a linear ODE with interrogate_kramer the model is linear-Gaussian, so solve_sim draws are exactly
    N(solve_mv mean, solve_mv var) marginally.  16,384 draws of one theta."""
    n = 16384
    pr = P.second_order_problem(1, n_steps=60, t_max=3.0, sigma=0.5, seed=0)
    X0 = np.repeat(pr["X0"], n, axis=0); th = np.repeat(pr["theta"], n, axis=0)
    x = _np(rb.solve_sim(np.array([3, 4], dtype=np.uint32), rb.models.second_order_sin, pr["W"], X0, 0.0, 3.0, 60,
                         rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=th))
    om, ov = orc.solve_mv(orc.MODELS["second_order_sin"], pr["W"], pr["X0"], 0.0, 3.0, 60, orc.interrogate_kramer,
                          (pr["Q"], pr["R"]), pr["theta"])
    assert np.array_equal(x[:, 0], X0)                      # row 0 is ode_init verbatim
    for t in (10, 30, 59, 60):
        sd = np.sqrt(np.maximum(np.diagonal(ov[0, t, 0]), 0))
        zmean = (x[:, t, 0].mean(axis=0) - om[0, t, 0]) / np.maximum(sd / np.sqrt(n), 1e-300)
        keep = sd > 1e-9 * np.abs(om[0, t, 0]).max()
        assert np.all(np.abs(zmean[keep]) < 5.0), (t, zmean)
        cov = np.cov(x[:, t, 0].T)
        scale = np.outer(sd, sd)[np.ix_(keep, keep)]
        assert np.max(np.abs(cov[np.ix_(keep, keep)] - ov[0, t, 0][np.ix_(keep, keep)]) / scale) < 0.06, t


def test_solve_sim_is_deterministic_and_sharding_invariant(rb):
    pr = P.fitz_problem(64, n_steps=80, t_max=4.0, seed=15)
    chk = _interr(rb, "chkrebtii")
    kw = dict(prior_pars=(pr["Q"], pr["R"]))
    key = np.array([11, 22], dtype=np.uint32)
    a = _np(rb.solve_sim(key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 4.0, 80, chk, theta=pr["theta"], **kw))
    b = _np(rb.solve_sim(key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 4.0, 80, chk, theta=pr["theta"], **kw))
    assert np.array_equal(a, b) and np.isfinite(a).all()
    # second half computed as its own shard with the global particle offset
    c = _np(rb.solve_sim(key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"][32:], 0.0, 4.0, 80, chk,
                         theta=pr["theta"][32:], _particle_offset=32, **kw))
    assert np.array_equal(a[32:], c)
    d = _np(rb.solve_sim(np.array([11, 23], dtype=np.uint32), rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 4.0,
                         80, chk, theta=pr["theta"], **kw))
    assert not np.array_equal(a, d)


def test_philox_normals_are_standard(rb):
    # terminal draw of a problem whose filter variance is known: x_N - mu_N = A z with A A^T = S_f[N]
    n = 1 << 20
    pr = P.second_order_problem(1, n_steps=1, t_max=0.5, sigma=2.0, seed=0)
    X0 = np.repeat(pr["X0"], n, axis=0); th = np.repeat(pr["theta"], n, axis=0)
    x = _np(rb.solve_sim(7, rb.models.second_order_sin, pr["W"], X0, 0.0, 0.5, 1, rb.interrogate.interrogate_kramer,
                         prior_pars=(pr["Q"], pr["R"]), theta=th))
    om, ov = orc.solve_mv(orc.MODELS["second_order_sin"], pr["W"], pr["X0"], 0.0, 0.5, 1, orc.interrogate_kramer,
                          (pr["Q"], pr["R"]), pr["theta"])
    r = x[:, 1, 0] - om[0, 1, 0]
    sd = np.sqrt(np.diagonal(ov[0, 1, 0]))
    import scipy.stats
    for j in (0, 1, 3):
        z = r[:, j] / sd[j]
        assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 0.02
        assert abs(np.mean(z ** 3)) < 0.05 and abs(np.mean(z ** 4) - 3) < 0.1
        assert scipy.stats.kstest(z, "norm").pvalue > 1e-3
        for cut, p_tail in ((3.0, 2.6998e-3), (4.0, 6.334e-5)):          # two-sided tail masses
            k = int(np.sum(np.abs(z) > cut))
            assert abs(k - n * p_tail) < 5 * np.sqrt(n * p_tail) + 1, (cut, k)


def test_solve_sim_draws_do_not_depend_on_the_lane_mapping(rb, monkeypatch):
    """One lane per theta and one lane per (theta, block) read the same Philox stream layout (rodeo_kernels.cuh), so
    the same key gives the same trajectories whichever kernel the host picks for the batch size (agreement to rounding:
    the two kernels are compiled separately)."""
    chk = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="standard")
    for name, pr in (("fitzhugh_nagumo", P.fitz_problem(40, n_steps=50, t_max=2.5, seed=61)),
                     ("lorenz63", P.lorenz_problem(23, n_steps=40, t_max=0.2, sigma=1.0, seed=62))):
        out = []
        monkeypatch.setenv("RODEO_SIM_SCHEDULE", "0")          # the full kernels (tests/test_gpu_schedule.py covers the schedule)
        for force in ("0", "1"):
            monkeypatch.setenv("RODEO_SIM_BLOCK_LANES", force)
            out.append(_np(rb.solve_sim(np.array([3, 4], dtype=np.uint32), getattr(rb.models, name), pr["W"], pr["X0"],
                                        0.0, pr["t_max"], pr["n_steps"], chk, prior_pars=(pr["Q"], pr["R"]),
                                        theta=pr["theta"])))
        assert np.isfinite(out[0]).all() and P.maxnorm_rel(out[0], out[1]) < 1e-9, name


# ---- full-size properties --------------------------------------------------------------------------------------------
def test_dalton_full_size_every_theta_against_the_c_oracle(rb):
    """BASELINE configs[1] exactly as bench.py times it: all 65,536 thetas of the launch against the C port (float64)
    and its long-double build, per-theta gate of tests/noise_floor.py."""
    import noise_floor as NF
    from oracle import c_port
    B = 65536
    pr = P.fitz_problem(B, seed=0)
    pr0 = P.fitz_problem(1, jitter=False)
    truth, _ = c_port.solve_mv("fitzhugh_nagumo", "kramer", pr0["W"], pr0["X0"], 0.0, 40.0, 800, pr0["Q"], pr0["R"],
                               pr0["theta"])
    ob = P.fitz_obs(pr, truth[0])
    got = _np(rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 40.0, 800,
                                  rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]),
                                  theta=pr["theta"], **ob))
    assert got.shape == (B,) and np.isfinite(got).all()
    ind = orc.obs_index(0.0, 40.0, 800, ob["obs_times"])
    cargs = ("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"], 0.0, 40.0, 800, pr["Q"], pr["R"], pr["theta"],
             ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"])
    want, exact = c_port.dalton(*cargs), c_port.dalton_ld(*cargs)
    st = NF.gate(got, want, exact)
    NF.report("dalton FN kramer N=800 B=65536 (bench workload)", st)
    assert st["ok"], st
    # batch-composition invariance: a theta's result does not depend on its neighbours
    sub = np.random.default_rng(0).choice(B, 1024, replace=False)
    again = _np(rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"][sub], 0.0, 40.0, 800,
                                    rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]),
                                    theta=pr["theta"][sub], **ob))
    assert np.array_equal(again, got[sub])


def test_solve_mv_full_size_matches_c_oracle_on_a_subset(rb):
    """BASELINE configs[0] at the throughput batch: 65,536 thetas x 800 steps = 10.1 GB of output."""
    import torch
    from oracle import c_port
    B = 65536
    pr = P.fitz_problem(B, seed=0)
    m, v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 40.0, 800,
                       rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])
    assert m.shape == (B, 801, 2, 3) and v.shape == (B, 801, 2, 3, 3)
    assert bool(torch.isfinite(m).all()) and bool(torch.isfinite(v).all())
    # row 0 is (ode_init, 0) verbatim (solve.py:295-301); variances are symmetric
    assert torch.equal(m[:, 0], torch.as_tensor(pr["X0"], device=m.device)) and not bool(v[:, 0].any())
    assert torch.equal(v, v.transpose(-1, -2))
    sub = np.sort(np.random.default_rng(1).choice(B, 256, replace=False))
    om, ov = c_port.solve_mv("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"][sub], 0.0, 40.0, 800, pr["Q"], pr["R"],
                             pr["theta"][sub])
    idx = torch.as_tensor(sub, device=m.device)
    assert P.maxnorm_rel(_np(m[idx]), om) < TOL and P.maxnorm_rel(_np(v[idx]), ov) < TOL
    # batch-composition invariance: the same thetas alone give bitwise the same rows
    m2, v2 = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"][sub], 0.0, 40.0, 800,
                         rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"][sub])
    assert torch.equal(m2, m[idx]) and torch.equal(v2, v[idx])


def test_solve_sim_full_size_shards_reproduce_the_whole(rb):
    """BASELINE configs[4], one GPU's share: 32,768 particles x solve_sim.  The 8-GPU job cuts the particle axis into
    contiguous shards; a shard run on its own (with its particle offset) must reproduce its rows of the whole."""
    import torch
    B = 32768
    pr = P.fitz_problem(B, seed=0)
    chk = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="standard")
    key = np.array([11, 12], dtype=np.uint32)
    args = (rb.models.fitzhugh_nagumo, pr["W"])
    x = rb.solve_sim(key, *args, pr["X0"], 0.0, 40.0, 800, chk, prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])
    assert x.shape == (B, 801, 2, 3) and bool(torch.isfinite(x).all())
    lo, hi = 3 * B // 8, 4 * B // 8                                  # rank 3 of 8
    xs = rb.solve_sim(key, *args, pr["X0"][lo:hi], 0.0, 40.0, 800, chk, prior_pars=(pr["Q"], pr["R"]),
                      theta=pr["theta"][lo:hi], _particle_offset=lo)
    assert P.maxnorm_rel(_np(xs), _np(x[lo:hi])) < 1e-9             # the two batch sizes may pick different lane mappings
    # FitzHugh-Nagumo trajectories live in a bounded limit cycle (|V| < 2.2, |R| < 1.3 for theta near (.2, .2, 3))
    assert float(x[:, :, :, 0].abs().max()) < 10.0


# ---- NVRTC user models ------------------------------------------------------------------------------------------------
def test_nvrtc_user_model_matches_builtin_and_oracle(rb):
    """A user right-hand side given as a CUDA string (the device-side analogue of passing any Python `ode_fun`) runs
    the same kernel templates: with the FitzHugh-Nagumo formula it must reproduce the built-in functor (whose
    Jacobian is analytic; the user model's comes from dual numbers, like jax.jacfwd) and the oracle."""
    user = rb.models.CudaOde("fitz_user", n_block=2, n_bstate=3, n_theta=3, rhs="""
        X V = x[0][0], R = x[1][0];
        f[0][0] = th[2] * (V - V * V * V / T(3) + R);
        f[1][0] = T(-1) / th[2] * (V - th[0] + th[1] * R);""")
    pr = P.fitz_problem(48, n_steps=200, t_max=10.0, seed=21)
    kr = rb.interrogate.interrogate_kramer
    m, v = rb.solve_mv(None, user, pr["W"], pr["X0"], 0.0, 10.0, 200, kr, prior_pars=(pr["Q"], pr["R"]),
                       theta=pr["theta"])
    om, ov = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                          (pr["Q"], pr["R"]), pr["theta"])
    assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(v), ov) < TOL
    ob = P.fitz_obs(pr, None, n_obs=11)
    a = rb.inference.dalton(None, user, pr["W"], pr["X0"], 0.0, 10.0, 200, kr, prior_pars=(pr["Q"], pr["R"]),
                            theta=pr["theta"], **ob)
    b = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200, kr,
                            prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"], **ob)
    assert ll_err(_np(a), _np(b)) < 2e-9
    f = rb.inference.fenrir(None, user, pr["W"], pr["X0"], 0.0, 10.0, 200, kr, prior_pars=(pr["Q"], pr["R"]),
                            theta=pr["theta"], **ob)
    fo = orc.fenrir(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                    (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert ll_err(_np(f), fo) < TOL
    # first_order_pad evaluates the user functor on the device as well
    W, init = rb.utils.first_order_pad(user, 2, 3)
    assert P.maxnorm_rel(_np(init(pr["x0"], 0.0, theta=pr["theta"])), pr["X0"]) < 1e-15


def test_nvrtc_reports_compile_errors(rb):
    bad = rb.models.CudaOde("broken", n_block=1, n_bstate=3, n_theta=1, rhs="f[0][0] = undefined_symbol;")
    with pytest.raises(Exception) as e:
        rb.solve_mv(None, bad, np.array([[[0.0, 1.0, 0.0]]]), np.zeros((1, 3)), 0.0, 1.0, 4,
                    rb.interrogate.interrogate_kramer, prior_pars=rb.prior.ibm_init(0.25, 3, [1.0]), theta=np.ones(1))
    assert "undefined_symbol" in str(e.value)


# ---- float32 instantiation -----------------------------------------------------------------------------------------
def test_float32_solve_mv_and_dalton(rb):
    """float32 kernels against the FLOAT64 oracle (i.e. against the truth, not against another float32 run), at
    BASELINE's float32 gate of 1e-5.  A plain float32 evaluation of the reference algorithm drifts 1.4e-5 (N=200) to
    4e-5 (N=800) from float64: the state x integrates its derivative over N steps and the update
    residual ~1e-3 is a difference of O(1) numbers.  The *_f32 kernels therefore carry the block means, the ODE
    evaluation, the residuals and the log-density sums in double (MeanOf<T>, rodeo_core.cuh) and only the
    covariances / gains -- the bulk of the arithmetic -- in float.  Measured: mean 5.9e-6, var 1.4e-6, dalton 7.5e-6."""
    import torch
    kr = rb.interrogate.interrogate_kramer
    for N, tm, tol_m in ((200, 10.0, 1e-5), (800, 40.0, 1e-5)):
        pr = P.fitz_problem(64, n_steps=N, t_max=tm, seed=31)
        th32, X32 = pr["theta"].astype(np.float32), pr["X0"].astype(np.float32)
        m, v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], X32, 0.0, tm, N, kr,
                           prior_pars=(pr["Q"], pr["R"]), theta=th32)
        assert m.dtype == torch.float32 and v.dtype == torch.float32
        om, ov = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], X32.astype(np.float64), 0.0, tm, N,
                              orc.interrogate_kramer, (pr["Q"], pr["R"]), th32.astype(np.float64))
        em, ev = P.maxnorm_rel(_np(m), om), P.maxnorm_rel(_np(v), ov)
        print(f"float32 N={N}: mean {em:.2e} var {ev:.2e}")
        assert em < tol_m and ev < 1e-5
    pr = P.fitz_problem(64, n_steps=200, t_max=10.0, seed=32)
    ob = P.fitz_obs(pr, None, n_obs=11)
    th32, X32 = pr["theta"].astype(np.float32), pr["X0"].astype(np.float32)
    ll = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], X32, 0.0, 10.0, 200, kr,
                             prior_pars=(pr["Q"], pr["R"]), theta=th32, **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], X32.astype(np.float64), 0.0, 10.0, 200,
                      orc.interrogate_kramer, (pr["Q"], pr["R"]), th32.astype(np.float64), ob["obs_data"],
                      ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    e = ll_err(_np(ll).astype(np.float64), want)
    print(f"float32 dalton: {e:.2e}")
    assert ll.dtype == torch.float32 and e < 1e-5


def test_float32_fenrir_and_solve_sim(rb):
    """The float32 fenrir log-likelihood and the float32 solve_sim draws (injected normals) against the FLOAT64 oracle,
    at BASELINE's float32 gate of 1e-5 (measured: fenrir 9.8e-8, draws 4.9e-6 -- means and residuals are carried in
    double, only covariances / gains / factors in float)."""
    import torch
    kr = rb.interrogate.interrogate_kramer
    pr = P.fitz_problem(64, n_steps=200, t_max=10.0, seed=33)
    ob = P.fitz_obs(pr, None, n_obs=11)
    th32, X32 = pr["theta"].astype(np.float32), pr["X0"].astype(np.float32)
    ll = rb.inference.fenrir(None, rb.models.fitzhugh_nagumo, pr["W"], X32, 0.0, 10.0, 200, kr,
                             prior_pars=(pr["Q"], pr["R"]), theta=th32, **ob)
    want = orc.fenrir(orc.MODELS["fitzhugh_nagumo"], pr["W"], X32.astype(np.float64), 0.0, 10.0, 200,
                      orc.interrogate_kramer, (pr["Q"], pr["R"]), th32.astype(np.float64), ob["obs_data"],
                      ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    e = ll_err(_np(ll).astype(np.float64), want)
    print(f"float32 fenrir: {e:.2e}")
    assert ll.dtype == torch.float32 and np.isfinite(_np(ll)).all() and e < 1e-5
    # solve_sim, float32, the same standard normals as the oracle (kramer: the full kernels)
    N = 120
    pr = P.fitz_problem(24, n_steps=N, t_max=6.0, seed=34)
    th32, X32 = pr["theta"].astype(np.float32), pr["X0"].astype(np.float32)
    zs = np.random.default_rng(8).standard_normal((24, N + 1, 2, 3))
    x = rb.solve_sim(0, rb.models.fitzhugh_nagumo, pr["W"], X32, 0.0, 6.0, N, kr, prior_pars=(pr["Q"], pr["R"]),
                     theta=th32, _z_smooth=zs.astype(np.float32))
    want = orc.solve_sim(orc.MODELS["fitzhugh_nagumo"], pr["W"], X32.astype(np.float64), 0.0, 6.0, N,
                         orc.interrogate_kramer, (pr["Q"], pr["R"]), th32.astype(np.float64),
                         z_smooth=zs.astype(np.float32).astype(np.float64), factor="ldl")
    ex = P.maxnorm_rel(_np(x).astype(np.float64), want)
    print(f"float32 solve_sim draws: {ex:.2e}")
    assert x.dtype == torch.float32 and ex < 1e-5


# ---- per-theta prior ---------------------------------------------------------------------------------------------------
def test_theta_dependent_prior_scale(rb):
    """sigma as part of theta (reference docs/examples/parameter.md:218-222: prior_pars rebuilt per theta under vmap):
    ibm_init with a (B, n_block) sigma gives R of shape (B, n_block, p, p)."""
    B = 48
    pr = P.fitz_problem(B, n_steps=200, t_max=10.0, seed=41)
    sig = 0.1 * np.exp(0.3 * np.random.default_rng(3).standard_normal((B, 2)))
    Q, Rb = rb.prior.ibm_init(10.0 / 200, 3, sig)
    assert Rb.shape == (B, 2, 3, 3)
    Qo = Q
    Ro = np.stack([orc.ibm_init(10.0 / 200, 3, sig[k])[1] for k in range(B)])
    assert np.allclose(Rb, Ro, rtol=1e-15)
    kr = rb.interrogate.interrogate_kramer
    m, v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200, kr, prior_pars=(Q, Rb),
                       theta=pr["theta"])
    om, ov = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                          (Qo, Ro), pr["theta"])
    assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(v), ov) < TOL
    ob = P.fitz_obs(pr, None, n_obs=11)
    ll = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200, kr,
                             prior_pars=(Q, Rb), theta=pr["theta"], **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                      (Qo, Ro), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert ll_err(_np(ll), want) < 2e-9
    # a per-theta prior that is NOT a multiple of one matrix takes the general (B, n_block, p, p) device-array path
    bad = Ro.copy(); bad[3, 0, 0, 1] *= 1.01; bad[3, 0, 1, 0] = bad[3, 0, 0, 1]
    m, v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200, kr, prior_pars=(Q, bad),
                       theta=pr["theta"])
    om, ov = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                          (Qo, bad), pr["theta"])
    assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(v), ov) < TOL
    ll = rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200, kr,
                             prior_pars=(Q, bad), theta=pr["theta"], **ob)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                      (Qo, bad), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert ll_err(_np(ll), want) < 2e-9
    fl = rb.inference.fenrir(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 10.0, 200, kr,
                             prior_pars=(Q, bad), theta=pr["theta"], **ob)
    want = orc.fenrir(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, orc.interrogate_kramer,
                      (Qo, bad), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert ll_err(_np(fl), want) < 2e-9


# ---- ragged batch sizes and out-of-bounds canaries (compute-sanitizer is closed on this pool) -----------------------------
@pytest.mark.parametrize("B", [1, 5, 33, 47])
def test_ragged_batches_and_output_canaries(rb, B):
    """Batch sizes that are not multiples of the warp's theta count: results must equal the corresponding rows of a
    larger batch, and nothing may be written outside the output / workspace buffers (guard words on both sides)."""
    import ctypes
    import torch
    from rodeo_b200 import _host, _lib
    N, tm = 37, 2.0                                         # N not a multiple of any segment length
    big = P.fitz_problem(64, n_steps=N, t_max=tm, seed=51)
    kr = rb.interrogate.interrogate_kramer
    ref_m, ref_v = rb.solve_mv(None, rb.models.fitzhugh_nagumo, big["W"], big["X0"], 0.0, tm, N, kr,
                               prior_pars=(big["Q"], big["R"]), theta=big["theta"])
    pb = _host.Problem(None, rb.models.fitzhugh_nagumo, big["W"], big["X0"][:B], 0.0, tm, N, kr,
                       (big["Q"], big["R"]), None, None, "standard", {"theta": big["theta"][:B]})
    G, SENT = 4096, -777.25

    def guarded(n):
        t = torch.full((n + 2 * G,), SENT, dtype=torch.float64, device="cuda")
        return t, t[G:G + n]

    nm, nv = B * (N + 1) * 6, B * (N + 1) * 18
    wsb = pb.lib.rodeo_b200_workspace_bytes(_lib.OP_SOLVE_MV, ctypes.byref(pb.c), 8)
    (gm, m), (gv, v), (gw, w) = guarded(nm), guarded(nv), guarded(max(wsb // 8, 1))
    rc = pb.fn("solve_mv")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R), _host.ptr(pb.x0),
                           _host.ptr(pb.theta), None, _host.ptr(m), _host.ptr(v), _host.ptr(w), wsb, pb.stream())
    _lib.check(rc, "solve_mv")
    torch.cuda.synchronize()
    for g, n in ((gm, nm), (gv, nv), (gw, max(wsb // 8, 1))):
        assert bool((g[:G] == SENT).all()) and bool((g[G + n:] == SENT).all()), "write outside the buffer"
    assert not bool((m == SENT).any()) and not bool((v == SENT).any()), "output not fully written"
    assert torch.equal(m.view(B, N + 1, 2, 3), ref_m[:B]) and torch.equal(v.view(B, N + 1, 2, 3, 3), ref_v[:B])

    # solve_sim (injected normals) and the two log-likelihoods on the same ragged batch vs the big batch
    rng = np.random.default_rng(0)
    zs = rng.standard_normal((64, N + 1, 2, 3))
    xs_big = rb.solve_sim(0, rb.models.fitzhugh_nagumo, big["W"], big["X0"], 0.0, tm, N, kr,
                          prior_pars=(big["Q"], big["R"]), theta=big["theta"], _z_smooth=zs)
    xs = rb.solve_sim(0, rb.models.fitzhugh_nagumo, big["W"], big["X0"][:B], 0.0, tm, N, kr,
                      prior_pars=(big["Q"], big["R"]), theta=big["theta"][:B], _z_smooth=zs[:B])
    assert torch.equal(xs.view(B, N + 1, 2, 3), xs_big[:B])
    ob = P.fitz_obs(big, None, n_obs=5)
    for fn in (rb.inference.dalton, rb.inference.fenrir):
        a = fn(None, rb.models.fitzhugh_nagumo, big["W"], big["X0"], 0.0, tm, N, kr, prior_pars=(big["Q"], big["R"]),
               theta=big["theta"], **ob)
        b_ = fn(None, rb.models.fitzhugh_nagumo, big["W"], big["X0"][:B], 0.0, tm, N, kr,
                prior_pars=(big["Q"], big["R"]), theta=big["theta"][:B], **ob)
        assert torch.equal(b_.view(B), a[:B])


def test_empty_batch(rb):
    """B = 0: every entry point returns correctly shaped empty tensors and launches nothing out of bounds."""
    pr = P.fitz_problem(4, n_steps=20, t_max=1.0, seed=2)
    ob = P.fitz_obs(pr, None, n_obs=3)
    kr = rb.interrogate.interrogate_kramer
    a = (None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"][:0], 0.0, 1.0, 20, kr)
    kw = dict(prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"][:0])
    m, v = rb.solve_mv(*a, **kw)
    assert m.shape == (0, 21, 2, 3) and v.shape == (0, 21, 2, 3, 3)
    assert rb.solve_sim(1, *a[1:], **kw).shape == (0, 21, 2, 3)
    assert rb.inference.dalton(*a, **kw, **ob).shape == (0,)
    assert rb.inference.fenrir(*a, **kw, **ob).shape == (0,)


# ---- data-adaptive solvers ---------------------------------------------------------------------------------
def test_dalton_data_adaptive_solvers(rb):
    """rodeo.inference.dalton.solve_mv / solve_sim (dalton.py:374-545): observations enter the forward filter."""
    for name, pr, obf in (("fitzhugh_nagumo", P.fitz_problem(40, n_steps=120, t_max=6.0, seed=61), P.fitz_obs),
                          ("second_order_sin", P.second_order_problem(24, n_steps=150, t_max=5.0, sigma=1.0, seed=61),
                           P.second_order_obs)):
        N, tm = pr["n_steps"], pr["t_max"]
        ob = obf(pr, None, n_obs=7) if name == "fitzhugh_nagumo" else obf(pr, n_obs=6)
        mdl, om_ = getattr(rb.models, name), orc.MODELS[name]
        kr = rb.interrogate.interrogate_kramer
        m, v = rb.inference.dalton_solve_mv(None, mdl, pr["W"], pr["X0"], 0.0, tm, N, kr, prior_pars=(pr["Q"], pr["R"]),
                                            theta=pr["theta"], **ob)
        om, ov = orc.dalton_solve_mv(om_, pr["W"], pr["X0"], 0.0, tm, N, orc.interrogate_kramer, (pr["Q"], pr["R"]),
                                     pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
        assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(v), ov) < TOL, name
        # the data must actually matter: differs from the plain solver
        m0, _ = rb.solve_mv(None, mdl, pr["W"], pr["X0"], 0.0, tm, N, kr, prior_pars=(pr["Q"], pr["R"]),
                            theta=pr["theta"])
        if name == "fitzhugh_nagumo":      # (the linear ODE is pinned by its interrogations; data barely move it)
            assert P.maxnorm_rel(_np(m), _np(m0)) > 1e-6
        rng = np.random.default_rng(1)
        zs = rng.standard_normal((pr["X0"].shape[0], N + 1) + pr["X0"].shape[1:])
        x = rb.inference.dalton_solve_sim(0, mdl, pr["W"], pr["X0"], 0.0, tm, N, kr, prior_pars=(pr["Q"], pr["R"]),
                                          theta=pr["theta"], _z_smooth=zs, **ob)
        ox = orc.dalton_solve_sim(om_, pr["W"], pr["X0"], 0.0, tm, N, orc.interrogate_kramer, (pr["Q"], pr["R"]),
                                  pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"],
                                  z_smooth=zs, factor="ldl")
        # singular smoothing covariances: a pivot that is zero in exact arithmetic is rounding noise e ~ 1e-16 * scale,
        # and its square root (~1e-8 * sqrt(scale)) multiplies a normal -- draws agree to ~1e-6, not 1e-10
        assert P.maxnorm_rel(_np(x), ox) < 2e-6, name


# ---- square-root Kalman family --------------------------------------------------------------------------------
def _sq(L):
    return L @ np.swapaxes(L, -1, -2)


@pytest.mark.parametrize("interr", ["kramer", "schober", "chkrebtii"])
def test_solve_mv_square_root(rb, interr):
    """kalman_type="square-root" (reference src/rodeo/kalmantv/square_root.py): means to 1e-10, variances compared as
    L L^T (QR sign ambiguity; reference tests/test_square_root.py:11-16 does the same), lower-triangular output."""
    pr = P.fitz_problem(40, n_steps=150, t_max=7.5, seed=71)
    Rh = np.linalg.cholesky(pr["R"])
    zi = np.random.default_rng(2).standard_normal((40, 150, 1, 2, 3))
    fn = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="square-root") if interr == "chkrebtii" \
        else getattr(rb.interrogate, "interrogate_" + interr)
    m, L = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 7.5, 150, fn, prior_pars=(pr["Q"], Rh),
                       kalman_type="square-root", theta=pr["theta"], _z_interr=zi)
    oi = {"kramer": orc.interrogate_kramer, "schober": orc.interrogate_schober,
          "chkrebtii": orc.interrogate_chkrebtii_sqrt}[interr]
    om, oL = orc.solve_mv_sqrt(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 7.5, 150, oi, (pr["Q"], Rh),
                               pr["theta"], z_interrogate=zi[:, :, 0])
    L = _np(L)
    assert P.maxnorm_rel(_np(m), om) < TOL
    assert P.maxnorm_rel(_sq(L), _sq(oL)) < 1e-9
    assert not np.triu(L, 1).any() and not L[:, 0].any()
    if interr == "kramer":      # and it is the same posterior as the covariance form
        m2, v2 = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 7.5, 150, fn,
                             prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])
        assert P.maxnorm_rel(_np(m), _np(m2)) < 1e-9 and P.maxnorm_rel(_sq(L), _np(v2)) < 1e-8


@pytest.mark.parametrize("N,tm", [(200, 10.0), (800, 40.0)])
def test_solve_mv_square_root_float32(rb, N, tm):
    """The square-root form is the numerically robust one for float32: float32 factors and QR, means
    carried in double, against the FLOAT64 oracle at BASELINE's float32 gate of 1e-5."""
    import torch
    pr = P.fitz_problem(32, n_steps=N, t_max=tm, seed=73)
    Rh = np.linalg.cholesky(pr["R"])
    th32, X32 = pr["theta"].astype(np.float32), pr["X0"].astype(np.float32)
    m, L = rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], X32, 0.0, tm, N, rb.interrogate.interrogate_kramer,
                       prior_pars=(pr["Q"], Rh), kalman_type="square-root", theta=th32)
    assert m.dtype == torch.float32 and L.dtype == torch.float32
    om, oL = orc.solve_mv_sqrt(orc.MODELS["fitzhugh_nagumo"], pr["W"], X32.astype(np.float64), 0.0, tm, N,
                               orc.interrogate_kramer, (pr["Q"], Rh), th32.astype(np.float64))
    L = _np(L).astype(np.float64)
    em, ev = P.maxnorm_rel(_np(m).astype(np.float64), om), P.maxnorm_rel(_sq(L), _sq(oL))
    print(f"float32 square-root solve_mv N={N}: mean {em:.2e} var (L L^T) {ev:.2e}")
    assert em < 1e-5 and ev < 1e-4
    assert not np.triu(L, 1).any()


def test_square_root_higher_order_docs_example(rb):
    # reference docs/examples/higher_order.md:104-127: sigma = .001, n_steps = 400, prior_chol = cholesky(prior_R)
    pr = P.second_order_problem(8, n_steps=400, sigma=0.001, seed=3)
    Rh = np.linalg.cholesky(pr["R"])
    m, L = rb.solve_mv(None, rb.models.second_order_sin, pr["W"], pr["X0"], 0.0, 10.0, 400,
                       rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], Rh), kalman_type="square-root",
                       theta=pr["theta"])
    om, oL = orc.solve_mv_sqrt(orc.MODELS["second_order_sin"], pr["W"], pr["X0"], 0.0, 10.0, 400,
                               orc.interrogate_kramer, (pr["Q"], Rh), pr["theta"])
    assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_sq(_np(L)), _sq(oL)) < 1e-8


# ---- device-resident pseudo-marginal MCMC ---------------------------------------------------------------------
def test_pseudo_marginal_random_walk_many_chains(rb):
    """The reference's pseudo-marginal RW-MH (src/rodeo/inference/pseudo_marginal.py) with solve_sim + chkrebtii inside
    (docs/examples/parameter.md:333-396), here for 256 chains at once without leaving the device."""
    import torch
    from rodeo_b200.inference import pseudo_marginal as pm
    N, tm, C = 100, 5.0, 256
    pr0 = P.fitz_problem(1, n_steps=N, t_max=tm, jitter=False)
    truth, _ = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr0["W"], pr0["X0"], 0.0, tm, N, orc.interrogate_kramer,
                            (pr0["Q"], pr0["R"]), pr0["theta"])
    ob = P.fitz_obs(pr0, truth[0], n_obs=6)
    Y = ob["obs_data"][:, :, 0]
    ind = orc.obs_index(0.0, tm, N, ob["obs_times"])
    # fused observation log-likelihood == the docs' fitz_loglik on Xt[obs_ind]
    Xt = torch.as_tensor(np.repeat(truth, 3, axis=0)).cuda()
    ll = _np(pm.gauss_obs_loglik(Xt, ind, Y, np.sqrt(0.005)))
    import scipy.stats
    want = scipy.stats.norm.logpdf(Y, loc=truth[0][ind, :, 0], scale=np.sqrt(0.005)).sum()
    assert np.allclose(ll, want, rtol=1e-12)

    chk = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="standard")
    W, fitz_init = rb.utils.first_order_pad(rb.models.fitzhugh_nagumo, 2, 3)

    def logpost(upars, key):                       # upars = (log a, log b, log c, V0, R0), one row per chain
        theta = torch.exp(upars[:, :3]); x0 = upars[:, 3:5]
        X0 = fitz_init(x0, 0.0, theta=theta)
        ll, xs = rb.solve_sim_loglik(key, rb.models.fitzhugh_nagumo, W, X0, 0.0, tm, N, chk,
                                      prior_pars=(pr0["Q"], pr0["R"]), theta=theta, obs_data=Y,
                                      obs_times=ob["obs_times"], noise_sd=np.sqrt(0.005), return_draws=True)
        return ll - 0.5 * (upars ** 2).sum(dim=1) / 100.0, xs

    start = np.tile(np.concatenate([np.log([0.2, 0.2, 3.0]), [-1.0, 1.0]]), (C, 1))
    alg = pm.normal_random_walk(logpost, sigma=np.array([0.02, 0.02, 0.01, 0.01, 0.01]))
    state = alg.init(start, rng_key=1)
    assert state.position.shape == (C, 5) and state.auxdata.shape == (C, N + 1, 2, 3)
    n_acc = torch.zeros(C, device="cuda")
    for it in range(30):
        state, info = alg.step(np.array([7, it], dtype=np.uint32), state)
        n_acc += info.is_accepted
        assert 0.0 <= float(info.acceptance_rate.min()) and float(info.acceptance_rate.max()) <= 1.0
    rate = float(n_acc.mean() / 30)
    assert torch.isfinite(state.logdensity).all() and torch.isfinite(state.position).all()
    assert 0.02 < rate < 0.98, rate
    # reproducible for the same keys
    s2 = alg.init(start, rng_key=1)
    for it in range(3):
        s2, _ = alg.step(np.array([7, it], dtype=np.uint32), s2)
    s3 = alg.init(start, rng_key=1)
    for it in range(3):
        s3, _ = alg.step(np.array([7, it], dtype=np.uint32), s3)
    assert torch.equal(s2.position, s3.position)


def test_solve_sim_loglik_fused_equals_the_two_calls(rb):
    """rodeo_b200_solve_sim_loglik_f64 (draw + Gaussian observation log-likelihood in one kernel, trajectories optional)
    == solve_sim followed by gauss_obs_loglik, for both lane mappings, with and without writing the draws, with an
    observation at t_min and one at t_max."""
    import os
    from rodeo_b200.inference import pseudo_marginal as pm
    N, tm, B = 160, 8.0, 77
    pr = P.fitz_problem(B, n_steps=N, t_max=tm, seed=37)
    ob = P.fitz_obs(pr, None, n_obs=9)
    Y = ob["obs_data"][:, :, 0]
    ind = orc.obs_index(0.0, tm, N, ob["obs_times"])
    assert ind[0] == 0 and ind[-1] == N
    chk = _interr(rb, "chkrebtii")
    key = np.array([3, 9], dtype=np.uint32)
    common = dict(prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])
    for lanes in ("0", "1"):
        os.environ["RODEO_SIM_BLOCK_LANES"] = lanes
        os.environ["RODEO_SIM_SCHEDULE"] = "0"
        try:
            x = rb.solve_sim(key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, tm, N, chk, **common)
            want = _np(pm.gauss_obs_loglik(x, ind, Y, 0.07))
            ll = rb.solve_sim_loglik(key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, tm, N, chk,
                                     obs_data=Y, obs_times=ob["obs_times"], noise_sd=0.07, **common)
            ll2, x2 = rb.solve_sim_loglik(key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, tm, N, chk,
                                          obs_data=Y, obs_times=ob["obs_times"], noise_sd=0.07, return_draws=True,
                                          **common)
        finally:
            del os.environ["RODEO_SIM_BLOCK_LANES"]
            del os.environ["RODEO_SIM_SCHEDULE"]
        assert np.array_equal(_np(x2), _np(x)) and np.array_equal(_np(ll2), _np(ll))
        assert ll_err(_np(ll), want) < 1e-12
        assert ll_err(want, orc.gauss_obs_loglik(_np(x), ind, Y, 0.07)) < 1e-12


def test_pseudo_marginal_chain_equals_the_oracle_chain(rb):
    """A short pseudo-marginal RW-MH chain (proposal kernel, fused solve_sim + observation log-likelihood kernel, accept
    kernel) with every random number injected -- proposal normals, acceptance uniforms, the chkrebtii and smoothing
    normals of each solve -- against the oracle's chain, which restates the reference's step
    (src/rodeo/inference/pseudo_marginal.py:452-483, 332-379) around the oracle's solve_sim."""
    import torch
    from rodeo_b200.inference import pseudo_marginal as pm
    N, tm, C, n_it = 60, 3.0, 48, 6
    pr0 = P.fitz_problem(1, n_steps=N, t_max=tm, jitter=False)
    truth, _ = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr0["W"], pr0["X0"], 0.0, tm, N, orc.interrogate_kramer,
                            (pr0["Q"], pr0["R"]), pr0["theta"])
    ob = P.fitz_obs(pr0, truth[0], n_obs=4)
    Y, sd = ob["obs_data"][:, :, 0], 0.2
    ind = orc.obs_index(0.0, tm, N, ob["obs_times"])
    rng = np.random.default_rng(11)
    zs = rng.standard_normal((n_it + 1, C, N + 1, 2, 3))
    zi = rng.standard_normal((n_it + 1, C, N, 1, 2, 3))
    zp = rng.standard_normal((n_it, C, 5))
    us = rng.uniform(size=(n_it, C))
    sigma = np.array([0.05, 0.05, 0.03, 0.02, 0.02])
    chk = _interr(rb, "chkrebtii")
    ochk = functools.partial(orc.interrogate_chkrebtii, factor="ldl")
    W, fitz_init = rb.utils.first_order_pad(rb.models.fitzhugh_nagumo, 2, 3)
    it = {"k": 0}

    def logpost(upars, key):                                   # CUDA: one fused kernel + the prior
        theta = torch.exp(upars[:, :3])
        X0 = fitz_init(upars[:, 3:5], 0.0, theta=theta)
        ll = rb.solve_sim_loglik(0, rb.models.fitzhugh_nagumo, W, X0, 0.0, tm, N, chk, prior_pars=(pr0["Q"], pr0["R"]),
                                 theta=theta, obs_data=Y, obs_times=ob["obs_times"], noise_sd=sd,
                                 _z_smooth=zs[it["k"]], _z_interr=zi[it["k"]])
        return ll - 0.5 * (upars ** 2).sum(dim=1) / 100.0, None

    def o_logpost(upars, k):
        theta = np.exp(upars[:, :3]); x0 = upars[:, 3:5]
        X0 = np.zeros((C, 2, 3)); X0[:, :, 0] = x0; X0[:, :, 1] = P.fitz_rhs(x0, theta)
        x = orc.solve_sim(orc.MODELS["fitzhugh_nagumo"], pr0["W"], X0, 0.0, tm, N, ochk, (pr0["Q"], pr0["R"]), theta,
                          z_smooth=zs[k], z_interrogate=zi[k][:, :, 0], factor="ldl")
        return orc.gauss_obs_loglik(x, ind, Y, sd) - 0.5 * (upars ** 2).sum(axis=1) / 100.0

    start = np.tile(np.concatenate([np.log([0.2, 0.2, 3.0]), [-1.0, 1.0]]), (C, 1)) \
        + 0.02 * rng.standard_normal((C, 5))
    alg = pm.normal_random_walk(logpost, sigma=sigma)
    state = alg.init(start, rng_key=0)
    o_pos, o_ld = start.copy(), o_logpost(start, 0)
    assert state.auxdata is None and ll_err(_np(state.logdensity), o_ld) < 1e-7
    n_acc = 0
    for k in range(n_it):
        it["k"] = k + 1
        state, info = alg.step(k, state, _z=zp[k], _u=us[k])
        o_pos, o_ld, o_acc, o_pa = orc.rwmh_step(o_pos, o_ld, lambda q: o_logpost(q, k + 1), sigma, zp[k], us[k])
        assert np.array_equal(_np(info.is_accepted), o_acc)
        assert np.allclose(_np(info.acceptance_rate), o_pa, rtol=1e-6, atol=1e-9)
        assert P.maxnorm_rel(_np(state.position), o_pos) < 1e-14 and ll_err(_np(state.logdensity), o_ld) < 1e-7
        n_acc += int(o_acc.sum())
    assert 0 < n_acc < n_it * C


def test_fenrir_solve_mv(rb):
    """rodeo.inference.fenrir.solve_mv (fenrir.py:404-457)."""
    pr = P.fitz_problem(40, n_steps=100, t_max=5.0, seed=81)
    for times in (np.linspace(0.0, 5.0, 6), np.array([0.4, 1.3, 2.0, 3.7])):
        ob = P.fitz_obs(pr, None, n_obs=len(times)); ob["obs_times"] = times
        m, v = rb.inference.fenrir_solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, 100,
                                            rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]),
                                            theta=pr["theta"], **ob)
        om, ov = orc.fenrir_solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, 100,
                                     orc.interrogate_kramer, (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"],
                                     ob["obs_times"], ob["obs_weight"], ob["obs_var"])
        assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(v), ov) < 1e-9


# ---- BASELINE configs at their full size, against the oracle on a subset ----------------------------
def test_fenrir_c4_full_size_subset_against_the_numpy_oracle(rb):
    """BASELINE configs[3]: second-order ODE, p = 4, N = 2,000, 16,384 thetas, fenrir with 11
    observations.  The whole batch runs on the GPU; a 64-theta subset is checked against the NumPy oracle at 1e-10."""
    B = 16384
    pr = P.second_order_problem(B)
    ob = P.second_order_obs(pr)
    got = _np(rb.inference.fenrir(None, rb.models.second_order_sin, pr["W"], pr["X0"], 0.0, 10.0, 2000,
                                  rb.interrogate.interrogate_kramer, prior_pars=(pr["Q"], pr["R"]),
                                  theta=pr["theta"], **ob))
    assert got.shape == (B,) and np.isfinite(got).all()
    sub = np.sort(np.random.default_rng(4).choice(B, 64, replace=False))
    want = orc.fenrir(orc.MODELS["second_order_sin"], pr["W"], pr["X0"][sub], 0.0, 10.0, 2000,
                      orc.interrogate_kramer, (pr["Q"], pr["R"]), pr["theta"][sub], ob["obs_data"], ob["obs_times"],
                      ob["obs_weight"], ob["obs_var"])
    e = ll_err(got[sub], want)
    print(f"C4 fenrir full size, 64-theta subset vs NumPy oracle: {e:.2e}")
    assert e < TOL


@pytest.mark.parametrize("lanes", ["0", "1", "schedule"])
def test_solve_sim_c5_injected_normals_subset(rb, monkeypatch, lanes):
    """BASELINE configs[4] (C5): FitzHugh-Nagumo solve_sim + interrogate_chkrebtii at N = 800, a 64-theta subset of one
    GPU's 32,768 particles with the SAME standard normals injected into kernel and oracle, for both lane mappings the
    host may pick.  With chkrebtii the measurement noise W S_p W^T keeps every filtered covariance full rank, but the
    smoothing covariance S_f - G (S_f Q^T)^T is formed by cancellation: its factor, hence the draw, is reproducible
    to ~1e-9 of the state scale between two float64 evaluations, not 1e-10.  "schedule": the covariance-schedule kernel
    the host picks by default for this interrogation (rodeo_sched.cuh); "0" / "1": the full kernels."""
    monkeypatch.setenv("RODEO_SIM_SCHEDULE", "1" if lanes == "schedule" else "0")
    monkeypatch.setenv("RODEO_SIM_BLOCK_LANES", "0" if lanes == "schedule" else lanes)
    pr_all = P.fitz_problem(32768, seed=0)
    sub = np.sort(np.random.default_rng(5).choice(32768, 64, replace=False))
    rng = np.random.default_rng(6)
    zs = rng.standard_normal((64, 801, 2, 3))
    zi = rng.standard_normal((64, 800, 1, 2, 3))
    x = rb.solve_sim(0, rb.models.fitzhugh_nagumo, pr_all["W"], pr_all["X0"][sub], 0.0, 40.0, 800,
                     _interr(rb, "chkrebtii"), prior_pars=(pr_all["Q"], pr_all["R"]), theta=pr_all["theta"][sub],
                     _z_smooth=zs, _z_interr=zi)
    want = orc.solve_sim(orc.MODELS["fitzhugh_nagumo"], pr_all["W"], pr_all["X0"][sub], 0.0, 40.0, 800,
                         functools.partial(orc.interrogate_chkrebtii, factor="ldl"), (pr_all["Q"], pr_all["R"]),
                         pr_all["theta"][sub], z_smooth=zs, z_interrogate=zi[:, :, 0], factor="ldl")
    e = P.maxnorm_rel(_np(x), want)
    print(f"C5 solve_sim+chkrebtii N=800, 64 thetas, injected normals, block lanes={lanes}: {e:.2e}")
    assert np.isfinite(_np(x)).all() and e < 1e-8


def test_lorenz_c3_nan_pattern_matches_the_oracle(rb):
    """BASELINE configs[2] (C3): Lorenz63 solve_sim + interrogate_chkrebtii with sigma = 5e7.  The reference algorithm
    itself overflows on this set-up (draws from N(mu_p, S_p) with S_p ~ 1e15 fed to a quadratic right-hand side), so
    the config is a throughput / NaN-propagation case.  With the same injected normals the kernel must
    propagate non-finite values exactly where the oracle does: row 0 = ode_init, identical NaN pattern in the draws,
    and the forward filter (solve_mv) turning non-finite at the same step for every theta."""
    B, N = 16, 4000
    pr = P.lorenz_problem(B, n_steps=N)
    rng = np.random.default_rng(8)
    zs = rng.standard_normal((B, N + 1, 3, 3))
    zi = rng.standard_normal((B, N, 1, 3, 3))
    chk = _interr(rb, "chkrebtii")
    ochk = functools.partial(orc.interrogate_chkrebtii, factor="ldl")
    x = _np(rb.solve_sim(0, rb.models.lorenz63, pr["W"], pr["X0"], 0.0, 20.0, N, chk, prior_pars=(pr["Q"], pr["R"]),
                         theta=pr["theta"], _z_smooth=zs, _z_interr=zi))
    with np.errstate(all="ignore"):
        want = orc.solve_sim(orc.MODELS["lorenz63"], pr["W"], pr["X0"], 0.0, 20.0, N, ochk, (pr["Q"], pr["R"]),
                             pr["theta"], z_smooth=zs, z_interrogate=zi[:, :, 0], factor="ldl")
        om, _ = orc.solve_mv(orc.MODELS["lorenz63"], pr["W"], pr["X0"], 0.0, 20.0, N, ochk, (pr["Q"], pr["R"]),
                             pr["theta"], z_interrogate=zi[:, :, 0])
    assert np.array_equal(x[:, 0], pr["X0"])
    assert np.array_equal(np.isfinite(x), np.isfinite(want))
    m, _ = rb.solve_mv(0, rb.models.lorenz63, pr["W"], pr["X0"], 0.0, 20.0, N, chk, prior_pars=(pr["Q"], pr["R"]),
                       theta=pr["theta"], _z_interr=zi)
    m = _np(m)
    bad_k, bad_o = ~np.isfinite(m).all(axis=(2, 3)), ~np.isfinite(om).all(axis=(2, 3))      # (B, N+1)
    print("C3 first non-finite solve_mv row per theta (kernel / oracle):", bad_k.argmax(1)[:8], bad_o.argmax(1)[:8],
          "fraction non-finite", bad_k.mean(), bad_o.mean())
    assert np.array_equal(bad_k, bad_o)
    fin = ~bad_o
    if fin.any():
        assert P.maxnorm_rel(m[fin], om[fin]) < 1e-6


# ---- ADVICE r1 ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("chunks", ["1", "2", "8"])
def test_host_entry_point_with_a_per_theta_prior_scale(rb, monkeypatch, chunks):
    """rodeo_b200_dalton_f64_host cuts the batch into chunks; every chunk must read ITS rows of the per-theta prior
    scale (a HOST array in the *_host wrappers).  Compared with the device-pointer entry point."""
    import ctypes
    import torch
    from rodeo_b200 import _host, _lib
    monkeypatch.setenv("RODEO_HOST_CHUNKS", chunks)
    B = 203
    pr = P.fitz_problem(B, n_steps=120, t_max=6.0, seed=17)
    ob = P.fitz_obs(pr, None, n_obs=7)
    sig = 0.1 * np.exp(0.3 * np.random.default_rng(2).standard_normal((B, 2)))
    Q, Rb = rb.prior.ibm_init(6.0 / 120, 3, sig)
    kr = rb.interrogate.interrogate_kramer
    want = _np(rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 6.0, 120, kr,
                                   prior_pars=(Q, Rb), theta=pr["theta"], **ob))
    pb = _host.Problem(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 6.0, 120, kr, (Q, Rb), None, None,
                       "standard", {"theta": pr["theta"]})
    pb.set_obs(ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    scale = np.ascontiguousarray(pb.r_scale_host, dtype=np.float64)
    c = _lib.RodeoProblem.from_buffer_copy(pb.c)
    c.prior_var_scale = scale.ctypes.data                     # HOST pointer for the *_host entry point
    out = np.full(B, np.nan)
    h = lambda a: _host.ptr(np.ascontiguousarray(a))
    x0, th = np.ascontiguousarray(pr["X0"]), np.ascontiguousarray(pr["theta"])
    y, D, Om = (np.ascontiguousarray(ob[k]) for k in ("obs_data", "obs_weight", "obs_var"))
    ind = np.ascontiguousarray(pb.obs_ind_host)
    rc = pb.lib.rodeo_b200_dalton_f64_host(ctypes.byref(c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                           _host.ptr(x0), _host.ptr(th), _host.ptr(ind), _host.ptr(y), _host.ptr(D),
                                           _host.ptr(Om), _host.ptr(out))
    _lib.check(rc, "dalton_host")
    assert np.array_equal(out, want)
    # and the oracle, with the per-theta prior built per theta
    Ro = np.stack([orc.ibm_init(6.0 / 120, 3, sig[k])[1] for k in range(B)])
    ow = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 6.0, 120, orc.interrogate_kramer,
                    (Q, Ro), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert ll_err(out, ow) < TOL


def test_basic_forwards_a_theta_dependent_prior(rb):
    """basic() must solve every theta with ITS prior variance (ibm_init with a (B, n_block) sigma)."""
    import torch
    B = 24
    pr = P.fitz_problem(B, n_steps=100, t_max=5.0, seed=23)
    ob = P.fitz_obs(pr, None, n_obs=6)
    sig = 0.1 * np.exp(0.4 * np.random.default_rng(4).standard_normal((B, 2)))
    Q, Rb = rb.prior.ibm_init(5.0 / 100, 3, sig)
    y = torch.as_tensor(ob["obs_data"], device="cuda")
    gauss = lambda obs_data, ode_data, **kw: torch.sum(-0.5 * (y[None, :, :, 0] - ode_data[:, :, :, 0]) ** 2 / 0.005,
                                                       dim=(1, 2))
    ll, Xt = rb.inference.basic(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, 100,
                                rb.interrogate.interrogate_kramer, prior_pars=(Q, Rb), obs_data=ob["obs_data"],
                                obs_times=ob["obs_times"], obs_loglik=gauss, theta=pr["theta"])
    Ro = np.stack([orc.ibm_init(5.0 / 100, 3, sig[k])[1] for k in range(B)])
    om, _ = orc.solve_mv(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, 100, orc.interrogate_kramer,
                         (Q, Ro), pr["theta"])
    assert P.maxnorm_rel(_np(Xt), om) < TOL
    ind = orc.obs_index(0.0, 5.0, 100, ob["obs_times"])
    want = np.sum(-0.5 * (ob["obs_data"][None, :, :, 0] - om[:, ind][:, :, :, 0]) ** 2 / 0.005, axis=(1, 2))
    assert ll_err(_np(ll), want) < 1e-9


def _perturbed_prior(pr, B, seed, rel=0.03, ab=0.005):
    """every theta gets its own dense (Q, R): neither shared nor a multiple of one matrix"""
    rq = np.random.default_rng(seed)
    nb, p = pr["Q"].shape[0], pr["Q"].shape[1]
    Qg, Rg = np.repeat(pr["Q"][None], B, 0).copy(), np.repeat(pr["R"][None], B, 0).copy()
    for i in range(B):
        for b in range(nb):
            Qg[i, b] = Qg[i, b] * (1.0 + rel * rq.standard_normal((p, p))) + ab * rq.standard_normal((p, p))
            L = np.linalg.cholesky(Rg[i, b])
            a = 0.3 * rq.standard_normal((p, p))
            Rg[i, b] = L @ (np.eye(p) + a @ a.T) @ L.T
            Rg[i, b] = 0.5 * (Rg[i, b] + Rg[i, b].T)
    return Qg, Rg


@pytest.mark.parametrize("name", ["lorenz63", "second_order_sin"])
def test_general_per_theta_prior_other_models(rb, name):
    """(B, n_block, p, p) priors on the three-block Lorenz63 and the p = 4 second-order ODE (ragged batch sizes) against
    the NumPy oracle: solve_mv, dalton, fenrir, and solve_sim on injected normals."""
    B = 37
    if name == "lorenz63":
        pr = P.lorenz_problem(B, n_steps=60, t_max=0.3, sigma=1.0, seed=61); N, tm = 60, 0.3
        obs_t = np.array([0.0, 0.1, 0.2, 0.3])
    else:
        pr = P.second_order_problem(B, n_steps=80, t_max=2.0, sigma=0.1, seed=62); N, tm = 80, 2.0
        obs_t = np.array([0.0, 0.5, 1.0, 1.5, 2.0])
    nb, p = pr["Q"].shape[0], pr["Q"].shape[1]
    # (Lorenz63 is chaotic: a gentler perturbation keeps the 60-step solve finite)
    Qg, Rg = _perturbed_prior(pr, B, 63, *((0.005, 1e-5) if name == "lorenz63" else (0.03, 0.005)))
    rng = np.random.default_rng(64)
    D = np.zeros((len(obs_t), nb, 1, p)); D[..., 0] = 1.0
    ob = dict(obs_data=rng.standard_normal((len(obs_t), nb, 1)), obs_times=obs_t, obs_weight=D,
              obs_var=np.full((len(obs_t), nb, 1, 1), 0.05))
    kr, fn, om_ = rb.interrogate.interrogate_kramer, getattr(rb.models, name), orc.MODELS[name]
    a = (None, fn, pr["W"], pr["X0"], 0.0, tm, N, kr)
    kw = dict(prior_pars=(Qg, Rg), theta=pr["theta"])
    oa = (om_, pr["W"], pr["X0"], 0.0, tm, N, orc.interrogate_kramer, (Qg, Rg), pr["theta"])
    oo = (ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    m, v = rb.solve_mv(*a, **kw)
    wm, wv = orc.solve_mv(*oa)
    assert P.maxnorm_rel(_np(m), wm) < TOL and P.maxnorm_rel(_np(v), wv) < 1e-9
    assert ll_err(_np(rb.inference.dalton(*a, **kw, **ob)), orc.dalton(*oa, *oo)) < 1e-9
    assert ll_err(_np(rb.inference.fenrir(*a, **kw, **ob)), orc.fenrir(*oa, *oo)) < 1e-9
    zs = rng.standard_normal((B, N + 1, nb, p))
    x = rb.solve_sim(0, *a[1:], **kw, _z_smooth=zs)
    wx = orc.solve_sim(*oa, z_smooth=zs, factor="ldl")
    assert P.maxnorm_rel(_np(x), wx) < 1e-6
    # an empty batch through the same path
    e = rb.solve_mv(None, fn, pr["W"], pr["X0"][:0], 0.0, tm, N, kr, prior_pars=(Qg[:0], Rg[:0]), theta=pr["theta"][:0])
    assert e[0].shape[0] == 0


# ---- dalton launch geometries ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("interr", ["kramer", "rodeo", "chkrebtii"])
def test_dalton_block_lane_kernel_is_bitwise_the_thread_per_filter_kernel(rb, monkeypatch, interr):
    """Small batches run dalton with one lane per (theta, filter, block) (dalton_bl_kernel), large ones with one thread
    per (theta, filter): the host chooses by batch size, so the two must agree BITWISE (per-block partial sums combined
    in block order by both), and each must match the oracle.  Ragged batch (not a multiple of the 16 thetas a warp
    carries), observation at t_min and on ordinary steps."""
    B, N = 203, 240
    pr = P.fitz_problem(B, n_steps=N, t_max=12.0, seed=29)
    ob = P.fitz_obs(pr, None, n_obs=13)
    zi = np.random.default_rng(3).standard_normal((B, N, 2, 2, 3)) if interr == "chkrebtii" else None
    out = {}
    for lanes in ("0", "1"):
        monkeypatch.setenv("RODEO_DALTON_BLOCK_LANES", lanes)
        out[lanes] = _np(rb.inference.dalton(0, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 12.0, N,
                                             _interr(rb, interr), prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                                             _z_interr=zi, **ob))
    assert np.array_equal(out["0"], out["1"])
    kw = dict(z_interrogate=zi) if zi is not None else {}
    oi = functools.partial(orc.interrogate_chkrebtii, factor="ldl") if interr == "chkrebtii" else ORC_INTERR[interr]
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 12.0, N, oi, (pr["Q"], pr["R"]),
                      pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"], **kw)
    e = ll_err(out["1"], want)
    print(f"dalton block lanes [{interr}] vs oracle: {e:.2e}")
    assert e < (1e-9 if interr == "chkrebtii" else TOL)


def test_dalton_block_lanes_three_blocks_lorenz(rb, monkeypatch):
    """n_block = 3: 10 thetas per warp, two idle lanes; a per-theta prior scale on top."""
    B, N = 37, 120
    pr = P.lorenz_problem(B, n_steps=N, t_max=0.6, sigma=50.0, seed=31)
    rng = np.random.default_rng(5)
    n_obs = 7
    obs_times = np.linspace(0.0, 0.6, n_obs)
    obs_data = (np.array([-12.0, -5.0, 38.0]) + rng.standard_normal((n_obs, 3)))[:, :, None]
    obs_weight = np.zeros((n_obs, 3, 1, 3)); obs_weight[..., 0] = 1.0
    obs_var = np.full((n_obs, 3, 1, 1), 0.5)
    sig = 50.0 * np.exp(0.2 * rng.standard_normal((B, 3)))
    Q, Rb = rb.prior.ibm_init(0.6 / N, 3, sig)
    out = {}
    for lanes in ("0", "1"):
        monkeypatch.setenv("RODEO_DALTON_BLOCK_LANES", lanes)
        out[lanes] = _np(rb.inference.dalton(None, rb.models.lorenz63, pr["W"], pr["X0"], 0.0, 0.6, N,
                                             rb.interrogate.interrogate_kramer, prior_pars=(Q, Rb), theta=pr["theta"],
                                             obs_data=obs_data, obs_times=obs_times, obs_weight=obs_weight,
                                             obs_var=obs_var))
    assert np.array_equal(out["0"], out["1"])
    Ro = np.stack([orc.ibm_init(0.6 / N, 3, sig[k])[1] for k in range(B)])
    want = orc.dalton(orc.MODELS["lorenz63"], pr["W"], pr["X0"], 0.0, 0.6, N, orc.interrogate_kramer, (Q, Ro),
                      pr["theta"], obs_data, obs_times, obs_weight, obs_var)
    assert ll_err(out["1"], want) < 1e-9
