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
    """The dalton log-likelihood is a difference of two sums of ~N terms z^2/S + log S whose residuals z carry
    ~1e-10 relative float64 rounding noise: the float64 oracle itself sits up to 4e-10 (N=800) from the exact
    value (tests/ld_reference.py, x87 longdouble).  So the gate is: the kernel must be as close to the exact value
    as the oracle is (within 1e-10, or 3x the oracle's own error), and kernel-vs-oracle must stay within that
    measured noise floor."""
    import ld_reference as L
    pr = P.fitz_problem(96, n_steps=N, t_max=t_max, seed=5)
    ob = _fitz_truth_obs(pr, n_obs)
    got, want = _dalton_pair(rb, pr, ob, "kramer")
    ind = orc.obs_index(0.0, t_max, N, ob["obs_times"])
    exact = L.dalton_ld(L.fitz_fun_ld, L.fitz_jac_ld, pr["W"], pr["X0"], 0.0, t_max, N, pr["Q"], pr["R"],
                        pr["theta"], ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"]).astype(np.float64)
    e_oracle, e_kernel = ll_err(want, exact), ll_err(got, exact)
    print(f"N={N}: |oracle-exact|={e_oracle:.2e} |kernel-exact|={e_kernel:.2e} |kernel-oracle|={ll_err(got, want):.2e}")
    assert got.shape == (96,)
    assert e_kernel <= max(TOL, 3 * e_oracle)
    assert ll_err(got, want) <= max(TOL, 4 * e_oracle)


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
