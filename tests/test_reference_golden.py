"""The reference's OWN SOURCE, executed over oracle/jaxshim (tests/golden/make_reference_golden.py), produced
tests/golden/reference_vectors.npz.  Here the oracle (CPU) and the CUDA path (GPU, through the C ABI) are checked
against those committed vectors: this is what pins the oracle to the reference's code rather than to a restatement.
/root/reference is not needed at test time."""
import functools
import os

import numpy as np
import pytest

import problems as P
from oracle import rodeo_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))

TOL = 1e-10          # BASELINE north_star: 1e-10 relative in float64


def prob(tag):
    pr = {k: G[f"{tag}_in_{k}"] for k in ("W", "X0", "theta", "Q", "R")}
    ob = {k: G[f"{tag}_in_{k}"] for k in ("obs_data", "obs_times", "obs_weight", "obs_var") if f"{tag}_in_{k}" in G.files}
    return pr, ob


def ll_err(a, b):
    return float(np.max(np.abs(np.asarray(a) - b) / np.maximum(1.0, np.abs(b))))


MODEL = {"fitz": "fitzhugh_nagumo", "fitzmid": "fitzhugh_nagumo", "readme": "fitzhugh_nagumo", "lorenz": "lorenz63",
         "so": "second_order_sin", "fitzN2": "fitzhugh_nagumo", "fitzN3": "fitzhugh_nagumo", "fitzpast": "fitzhugh_nagumo", "fitzbobs2": "fitzhugh_nagumo", "hes1": "hes1", "seirah": "seirah"}


def grid(tag):
    _, t_max, n_steps = G[f"{tag}_in_grid"]
    return MODEL[tag], float(t_max), int(n_steps)


def oargs(tag, interr=orc.interrogate_kramer):
    model, t_max, N = grid(tag)
    pr, ob = prob(tag)
    a = (orc.MODELS[model], pr["W"], pr["X0"], 0.0, t_max, N, interr, (pr["Q"], pr["R"]), pr["theta"])
    o = (ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"]) if ob else None
    return a, o


# ---------------------------------------------------------------------------------------------------------------------
# CPU: oracle == reference source
# ---------------------------------------------------------------------------------------------------------------------
def test_inputs_are_what_the_reference_helpers_build():
    """tests/problems.py builds W, X0, Q, R with local closed forms; the reference's first_order_pad / ibm_init
    (utils.py:80-102, prior/ibm.py:65-88) must give the same arrays, and so must the oracle's."""
    pr, _ = prob("fitz")
    assert np.array_equal(G["fitz_ref_W"], pr["W"])
    assert P.maxnorm_rel(pr["X0"], G["fitz_ref_X0"]) < 1e-15
    assert P.maxnorm_rel(pr["Q"], G["fitz_ref_Q"]) < 1e-15 and P.maxnorm_rel(pr["R"], G["fitz_ref_R"]) < 1e-14
    Q, R = orc.ibm_init(3.0 / 60, 3, np.array([0.1, 0.1]))
    assert P.maxnorm_rel(Q, G["fitz_ref_Q"]) < 1e-15 and P.maxnorm_rel(R, G["fitz_ref_R"]) < 1e-14


@pytest.mark.parametrize("name", ["kramer", "schober", "rodeo"])
def test_oracle_solve_mv_fitz(name):
    a, _ = oargs("fitz", getattr(orc, "interrogate_" + name))
    m, v = orc.solve_mv(*a)
    assert P.maxnorm_rel(m, G[f"fitz_{name}_mean"]) < 1e-12
    assert P.maxnorm_rel(v, G[f"fitz_{name}_var"]) < 1e-12


def test_oracle_solve_mv_chkrebtii_same_normals():
    a, _ = oargs("fitz", orc.interrogate_chkrebtii)
    m, v = orc.solve_mv(*a, z_interrogate=G["fitz_chkrebtii_z"])
    assert P.maxnorm_rel(m, G["fitz_chkrebtii_mean"]) < 1e-11
    assert P.maxnorm_rel(v, G["fitz_chkrebtii_var"]) < 1e-11


def test_oracle_solve_sim_same_normals_svd_factor():
    a, _ = oargs("fitz")
    x = orc.solve_sim(*a, z_smooth=G["fitz_sim_z"], factor="svd")
    assert P.maxnorm_rel(x, G["fitz_sim_x"]) < 1e-10


@pytest.mark.parametrize("tag", ["fitz", "fitzmid", "fitzpast", "readme", "so", "fitzN2", "fitzN3", "fitzbobs2"])
def test_oracle_dalton_fenrir(tag):
    a, o = oargs(tag)
    assert ll_err(orc.fenrir(*a, *o), G[f"{tag}_fenrir"]) < 1e-11
    if f"{tag}_dalton" in G.files:
        assert ll_err(orc.dalton(*a, *o), G[f"{tag}_dalton"]) < 1e-11


def test_oracle_basic_and_data_adaptive_solvers():
    a, o = oargs("fitz")
    gauss = lambda y, x, th: np.sum(-0.5 * (y[None, :, :, 0] - x[:, :, :, 0]) ** 2 / 0.005
                                    - 0.5 * np.log(2 * np.pi * 0.005), axis=(1, 2))
    ll, Xt = orc.basic(*a, o[0], o[1], gauss)
    assert ll_err(ll, G["fitz_basic"]) < 1e-11 and P.maxnorm_rel(Xt, G["fitz_basic_Xt"]) < 1e-12
    m, v = orc.dalton_solve_mv(*a, *o)
    assert P.maxnorm_rel(m, G["fitz_dalton_mean"]) < 1e-12 and P.maxnorm_rel(v, G["fitz_dalton_var"]) < 1e-12
    m, v = orc.fenrir_solve_mv(*a, *o)
    assert P.maxnorm_rel(m, G["fitz_fenrir_mean"]) < 1e-11 and P.maxnorm_rel(v, G["fitz_fenrir_var"]) < 1e-10


def test_oracle_square_root_family():
    a, _ = oargs("fitz")
    pr, _ = prob("fitz")
    a = a[:7] + ((pr["Q"], np.linalg.cholesky(pr["R"])),) + a[8:]
    m, L = orc.solve_mv_sqrt(*a)
    assert P.maxnorm_rel(m, G["fitz_sqrt_mean"]) < 1e-11
    assert P.maxnorm_rel(L @ np.swapaxes(L, -1, -2), G["fitz_sqrt_var"]) < 1e-10


@pytest.mark.parametrize("tag", ["readme", "lorenz", "so", "fitzN2", "fitzN3", "hes1", "seirah"])
def test_oracle_solve_mv_other_models(tag):
    a, _ = oargs(tag)
    m, v = orc.solve_mv(*a)
    assert P.maxnorm_rel(m, G[f"{tag}_mean"]) < 1e-11
    assert P.maxnorm_rel(v, G[f"{tag}_var"]) < 1e-10


def test_oracle_theta_dependent_prior():
    a, o = oargs("fitz")
    QR = [orc.ibm_init(3.0 / 60, 3, sg) for sg in G["fitzsig_in_sigma"]]
    Q, R = QR[0][0], np.stack([r for _, r in QR])                       # Q shared, R (B, nb, p, p)
    a = a[:7] + ((Q, R),) + a[8:]
    m, v = orc.solve_mv(*a)
    assert P.maxnorm_rel(m, G["fitzsig_mean"]) < 1e-12 and P.maxnorm_rel(v, G["fitzsig_var"]) < 1e-12
    assert ll_err(orc.dalton(*a, *o), G["fitzsig_dalton"]) < 1e-11


def test_oracle_general_per_theta_prior():
    """every theta brings its own dense (Q, R) -- neither shared nor a multiple of a shared matrix"""
    a, o = oargs("fitz")
    a = a[:7] + ((G["fitzqr_in_Q"], G["fitzqr_in_R"]),) + a[8:]
    m, v = orc.solve_mv(*a)
    assert P.maxnorm_rel(m, G["fitzqr_mean"]) < 1e-12 and P.maxnorm_rel(v, G["fitzqr_var"]) < 1e-12
    assert ll_err(orc.dalton(*a, *o), G["fitzqr_dalton"]) < 1e-11
    assert ll_err(orc.fenrir(*a, *o), G["fitzqr_fenrir"]) < 1e-11


def pair1b():
    pr, _ = prob("fitz")
    ob = {k: G["pair1b_in_" + k] for k in ("obs_data", "obs_times", "obs_weight", "obs_var")}
    return {k: G["pair1b_in_" + k] for k in ("W", "X0", "Q", "R")}, pr["theta"], ob


def test_oracle_two_measurement_rows_per_block():
    """n_bmeas = 2: one block holding two variables, ode_weight (1, 2, 6)"""
    pr, theta, ob = pair1b()
    a = (orc.MODELS["pair_one_block"], pr["W"], pr["X0"], 0.0, 3.0, 60, orc.interrogate_kramer, (pr["Q"], pr["R"]), theta)
    o = (ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    m, v = orc.solve_mv(*a)
    assert P.maxnorm_rel(m, G["pair1b_mean"]) < 1e-12 and P.maxnorm_rel(v, G["pair1b_var"]) < 1e-12
    assert ll_err(orc.dalton(*a, *o), G["pair1b_dalton"]) < 1e-11
    assert ll_err(orc.fenrir(*a, *o), G["pair1b_fenrir"]) < 1e-11


def test_oracle_magi_logdens():
    pr, _ = prob("fitz")
    pr2, _ = prob("so")
    assert ll_err(orc.magi_logdens(G["magi_fitz_state"], 1, (pr["Q"], pr["R"])), G["magi_fitz"]) / 1e8 < 1e-12
    assert ll_err(orc.magi_logdens(G["magi_so_state"], 1, (pr2["Q"], pr2["R"])), G["magi_so"]) / 1e8 < 1e-12
    assert np.max(np.abs(orc.magi_logdens(G["magi_fitz_state"], 1, (pr["Q"], pr["R"])) / G["magi_fitz"] - 1)) < 1e-12
    # n_active >= 2: the reference's recursion is rounding-dominated (see make_reference_golden.py): two float64
    # evaluations of its formulas differ by tens of percent -- same order of magnitude only
    got = orc.magi_logdens(G["magi_so_state"], 3, (pr2["Q"], pr2["R"]))
    assert np.all(np.isfinite(got)) and np.all((got / G["magi_so_illcond"] > 1 / 3) & (got / G["magi_so_illcond"] < 3))


def test_oracle_kalman_primitives():
    k = {n[7:]: G[n] for n in G.files if n.startswith("kal_in_")}
    pm, pS = orc.predict(k["mu"], k["S"], k["c"], k["Qm"], k["Rm"])
    um, uS = orc.update(pm, pS, k["xm"], k["d"], k["Wm"], k["Vm"])
    fm, fS = orc.forecast(pm, pS, k["d"], k["Wm"], k["Vm"])
    sm, sS = orc.smooth_mv(k["mun"], k["Sn"], um, uS, pm, pS, k["Qm"])
    ssm, ssS = orc.smooth_sim(k["xn"], um, uS, pm, pS, k["Qm"])
    cA, cb, cV = orc.smooth_cond(um, uS, pm, pS, k["Qm"])
    lp = orc.multivariate_normal_logpdf(k["xm"], fm, fS)
    for name, val in zip(("pm", "pS", "um", "uS", "fm", "fS", "sm", "sS", "ssm", "ssS", "cA", "cb", "cV", "lp"),
                         (pm, pS, um, uS, fm, fS, sm, sS, ssm, ssS, cA, cb, cV, lp)):
        assert P.maxnorm_rel(val, G["kal_" + name]) < 1e-12, name
    lp = orc.multivariate_normal_logpdf(G["lpcut_in_x"], np.zeros((5, 2)), G["lpcut_in_cov"])
    assert np.allclose(lp, G["lpcut"], rtol=1e-14, atol=0)


# ---------------------------------------------------------------------------------------------------------------------
# GPU: CUDA path (drop-in Python API -> C ABI -> kernels) == reference source
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rb():
    import rodeo_b200
    from rodeo_b200 import _lib
    _lib.load()
    return rodeo_b200


def _np(t):
    return t.detach().cpu().numpy()


def gargs(rb, tag, interr=None):
    model, t_max, N = grid(tag)
    pr, ob = prob(tag)
    interr = interr or rb.interrogate.interrogate_kramer
    a = (None, getattr(rb.models, model), pr["W"], pr["X0"], 0.0, t_max, N, interr)
    return a, dict(prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"]), ob


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kramer", "schober", "rodeo"])
def test_cuda_solve_mv_fitz(rb, name):
    a, kw, _ = gargs(rb, "fitz", getattr(rb.interrogate, "interrogate_" + name))
    m, v = rb.solve_mv(*a, **kw)
    assert P.maxnorm_rel(_np(m), G[f"fitz_{name}_mean"]) < TOL
    assert P.maxnorm_rel(_np(v), G[f"fitz_{name}_var"]) < TOL


@pytest.mark.gpu
def test_cuda_solve_mv_chkrebtii_same_normals(rb):
    chk = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="standard")
    a, kw, _ = gargs(rb, "fitz", chk)
    m, v = rb.solve_mv(0, *a[1:], **kw, _z_interr=G["fitz_chkrebtii_z"])
    assert P.maxnorm_rel(_np(m), G["fitz_chkrebtii_mean"]) < TOL
    assert P.maxnorm_rel(_np(v), G["fitz_chkrebtii_var"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["fitz", "fitzmid", "fitzpast", "readme", "so", "fitzN2", "fitzN3", "fitzbobs2"])
def test_cuda_dalton_fenrir(rb, tag):
    a, kw, ob = gargs(rb, tag)
    assert ll_err(_np(rb.inference.fenrir(*a, **kw, **ob)), G[f"{tag}_fenrir"]) < TOL
    got = _np(rb.inference.dalton(*a, **kw, **ob))
    if tag != "readme":
        assert ll_err(got, G[f"{tag}_dalton"]) < TOL
        return
    # README walkthrough (N = 800 = BASELINE configs[0]): the sum sits on its float64 noise floor, so the kernel and the
    # reference's own value are both judged against the long-double evaluation (tests/noise_floor.py)
    import noise_floor as NF
    from oracle import c_port
    model, t_max, N = grid(tag)
    pr, _ = prob(tag)
    ind = orc.obs_index(0.0, t_max, N, ob["obs_times"])
    exact = c_port.dalton_ld(model, "kramer", pr["W"], pr["X0"], 0.0, t_max, N, pr["Q"], pr["R"], pr["theta"],
                             ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"])
    st = NF.gate(got, G[f"{tag}_dalton"], exact)
    NF.report("dalton README golden (reference source over the jax stand-in)", st)
    assert st["ok"], st


@pytest.mark.gpu
def test_cuda_basic_and_data_adaptive_solvers(rb):
    import torch
    a, kw, ob = gargs(rb, "fitz")
    y = torch.as_tensor(ob["obs_data"], device="cuda")

    def gauss(obs_data, ode_data, **params):
        return torch.sum(-0.5 * (y[None, :, :, 0] - ode_data[:, :, :, 0]) ** 2 / 0.005
                         - 0.5 * np.log(2 * np.pi * 0.005), dim=(1, 2))
    ll, Xt = rb.inference.basic(*a, **kw, obs_data=ob["obs_data"], obs_times=ob["obs_times"], obs_loglik=gauss)
    assert ll_err(_np(ll), G["fitz_basic"]) < TOL and P.maxnorm_rel(_np(Xt), G["fitz_basic_Xt"]) < TOL
    m, v = rb.inference.dalton_solve_mv(*a, **kw, **ob)
    assert P.maxnorm_rel(_np(m), G["fitz_dalton_mean"]) < TOL and P.maxnorm_rel(_np(v), G["fitz_dalton_var"]) < TOL
    m, v = rb.inference.fenrir_solve_mv(*a, **kw, **ob)
    assert P.maxnorm_rel(_np(m), G["fitz_fenrir_mean"]) < TOL and P.maxnorm_rel(_np(v), G["fitz_fenrir_var"]) < 1e-9


@pytest.mark.gpu
def test_cuda_square_root_family(rb):
    a, kw, _ = gargs(rb, "fitz")
    pr, _ = prob("fitz")
    m, L = rb.solve_mv(*a, prior_pars=(pr["Q"], np.linalg.cholesky(pr["R"])), theta=pr["theta"],
                       kalman_type="square-root")
    L = _np(L)
    assert P.maxnorm_rel(_np(m), G["fitz_sqrt_mean"]) < TOL
    assert P.maxnorm_rel(L @ np.swapaxes(L, -1, -2), G["fitz_sqrt_var"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["readme", "lorenz", "so", "fitzN2", "fitzN3", "hes1", "seirah"])
def test_cuda_solve_mv_other_models(rb, tag):
    a, kw, _ = gargs(rb, tag)
    m, v = rb.solve_mv(*a, **kw)
    assert P.maxnorm_rel(_np(m), G[f"{tag}_mean"]) < TOL
    assert P.maxnorm_rel(_np(v), G[f"{tag}_var"]) < (1e-9 if tag in ("lorenz", "seirah") else TOL)


@pytest.mark.gpu
def test_cuda_theta_dependent_prior(rb):
    """sigma part of theta: rb.prior.ibm_init with a (B, n_block) sigma (the kernels carry it as a per-(theta, block)
    scale of the shared prior variance) against the reference's ibm_init called once per theta."""
    a, kw, ob = gargs(rb, "fitz")
    kw = dict(kw, prior_pars=rb.prior.ibm_init(3.0 / 60, 3, G["fitzsig_in_sigma"]))
    m, v = rb.solve_mv(*a, **kw)
    assert P.maxnorm_rel(_np(m), G["fitzsig_mean"]) < TOL and P.maxnorm_rel(_np(v), G["fitzsig_var"]) < TOL
    assert ll_err(_np(rb.inference.dalton(*a, **kw, **ob)), G["fitzsig_dalton"]) < TOL


@pytest.mark.gpu
def test_cuda_general_per_theta_prior(rb):
    """prior_pars = (Q, R) of shape (B, n_block, p, p), arbitrary per theta (what the reference accepts under vmap,
    docs/examples/parameter.md:218-236): device arrays read per thread (RodeoProblem.prior_batched)."""
    a, kw, ob = gargs(rb, "fitz")
    kw = dict(kw, prior_pars=(G["fitzqr_in_Q"], G["fitzqr_in_R"]))
    m, v = rb.solve_mv(*a, **kw)
    assert P.maxnorm_rel(_np(m), G["fitzqr_mean"]) < TOL and P.maxnorm_rel(_np(v), G["fitzqr_var"]) < TOL
    assert ll_err(_np(rb.inference.dalton(*a, **kw, **ob)), G["fitzqr_dalton"]) < TOL
    assert ll_err(_np(rb.inference.fenrir(*a, **kw, **ob)), G["fitzqr_fenrir"]) < TOL
    # the same call with the shared prior repeated per theta takes the batched kernels and must reproduce the shared ones
    pr, _ = prob("fitz")
    Qp = np.repeat(pr["Q"][None], 3, 0).copy(); Qp[1, 0, 0, 1] += 1e-3         # not a multiple: forces the batched path
    Rp = np.repeat(pr["R"][None], 3, 0)
    m2, v2 = rb.solve_mv(*a, **dict(kw, prior_pars=(Qp, Rp)))
    m0, v0 = rb.solve_mv(*a, **dict(kw, prior_pars=(pr["Q"], pr["R"])))
    keep = [0, 2]                                                               # thetas whose prior was not perturbed
    assert P.maxnorm_rel(_np(m2)[keep], _np(m0)[keep]) < 1e-12 and P.maxnorm_rel(_np(v2)[keep], _np(v0)[keep]) < 1e-12
    x2 = rb.solve_sim(0, *a[1:], **dict(kw, prior_pars=(Qp, Rp)))
    x0 = rb.solve_sim(0, *a[1:], **dict(kw, prior_pars=(pr["Q"], pr["R"])))
    assert np.all(np.isfinite(_np(x2))) and P.maxnorm_rel(_np(x2)[keep], _np(x0)[keep]) < 1e-7


@pytest.mark.gpu
def test_cuda_two_measurement_rows_per_block(rb):
    """n_bmeas = 2 (models.pair_one_block, the dense instantiation with a 2-row measurement update and, on observation
    steps, the stacked 3-row one) against the reference source's values."""
    pr, theta, ob = pair1b()
    kr = rb.interrogate.interrogate_kramer
    a = (None, rb.models.pair_one_block, pr["W"], pr["X0"], 0.0, 3.0, 60, kr)
    kw = dict(prior_pars=(pr["Q"], pr["R"]), theta=theta)
    m, v = rb.solve_mv(*a, **kw)
    assert P.maxnorm_rel(_np(m), G["pair1b_mean"]) < TOL and P.maxnorm_rel(_np(v), G["pair1b_var"]) < TOL
    assert ll_err(_np(rb.inference.dalton(*a, **kw, **ob)), G["pair1b_dalton"]) < TOL
    assert ll_err(_np(rb.inference.fenrir(*a, **kw, **ob)), G["pair1b_fenrir"]) < TOL
    for name in ("schober", "rodeo"):            # the other interrogations of the 2-row update, against the oracle
        it = getattr(rb.interrogate, "interrogate_" + name)
        m, v = rb.solve_mv(None, *a[1:7], it, **kw)
        om, ov = orc.solve_mv(orc.MODELS["pair_one_block"], pr["W"], pr["X0"], 0.0, 3.0, 60,
                              getattr(orc, "interrogate_" + name), (pr["Q"], pr["R"]), theta)
        assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(v), ov) < TOL


@pytest.mark.gpu
def test_cuda_magi_logdens(rb):
    """rodeo.inference.magi_logdens with the user's own ode_expand (plain NumPy here), batched over trajectories and
    un-batched, against the reference's values."""
    pr, _ = prob("fitz")

    def expand(U, theta):                                # (B, N+1, nb, 1) -> (B, N+1, nb, 3): (x, f(x, theta), 0)
        x = U[..., 0]
        f = np.stack([P.fitz_rhs(x[:, n], theta) for n in range(x.shape[1])], axis=1)
        return np.concatenate([U, f[..., None], np.zeros_like(U)], axis=-1)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) / b - 1)))
    got = rb.inference.magi_logdens(G["magi_fitz_in_U"], expand, 1, (pr["Q"], pr["R"]), "standard", theta=pr["theta"])
    assert got.shape == (3,) and rel(_np(got), G["magi_fitz"]) < TOL
    one = rb.inference.magi_logdens(G["magi_fitz_in_U"][1], lambda U, theta: expand(U[None], theta[None])[0], 1,
                                    (pr["Q"], pr["R"]), theta=pr["theta"][1])
    assert one.dim() == 0 and rel(float(one), G["magi_fitz"][1]) < TOL
    pr2, _ = prob("so")
    got = rb.inference.magi_logdens(None, lambda U, **kw: G["magi_so_state"], 1, (pr2["Q"], pr2["R"]))
    assert rel(_np(got), G["magi_so"]) < TOL
    # n_active >= 2: rounding-dominated in the reference itself (float64 evaluations of the same formulas differ by a
    # factor of 2..4 on these inputs); the kernel (packed symmetric variances) must stay finite
    for na, state, want, prior in ((2, "magi_fitz_state", "magi_fitz_illcond", (pr["Q"], pr["R"])),
                                  (3, "magi_so_state", "magi_so_illcond", (pr2["Q"], pr2["R"]))):
        got = _np(rb.inference.magi_logdens(None, lambda U, **kw: G[state], na, prior))
        assert np.all(np.isfinite(got)) and np.all(got < 0) and np.all(np.isfinite(G[want]))
    with pytest.raises(NotImplementedError):
        rb.inference.magi_logdens(None, lambda U, **kw: G["magi_so_state"], 3, (pr2["Q"], pr2["R"]), "square-root")


@pytest.mark.gpu
def test_cuda_first_order_pad_and_ibm_init(rb):
    pr, _ = prob("fitz")
    W, init = rb.utils.first_order_pad(rb.models.fitzhugh_nagumo, 2, 3)
    assert np.array_equal(np.asarray(W), G["fitz_ref_W"])
    X0 = init(pr["X0"][:, :, 0], 0.0, theta=pr["theta"])
    assert P.maxnorm_rel(_np(X0), G["fitz_ref_X0"]) < 1e-15
    Q, R = rb.prior.ibm_init(3.0 / 60, 3, np.array([0.1, 0.1]))
    assert P.maxnorm_rel(np.asarray(Q), G["fitz_ref_Q"]) < 1e-15 and P.maxnorm_rel(np.asarray(R), G["fitz_ref_R"]) < 1e-14


@pytest.mark.gpu
def test_cuda_kalman_primitives(rb):
    ktv = rb.kalmantv.standard
    k = {n[7:]: G[n] for n in G.files if n.startswith("kal_in_")}
    got = {n: [] for n in ("pm", "pS", "um", "uS", "fm", "fS", "sm", "sS", "ssm", "ssS", "cA", "cb", "cV", "lp")}
    for i in range(k["mu"].shape[0]):
        g = {n: v[i] for n, v in k.items()}
        pm, pS = ktv.predict(mean_state_past=g["mu"], var_state_past=g["S"], mean_state=g["c"], wgt_state=g["Qm"],
                             var_state=g["Rm"])
        um, uS = ktv.update(mean_state_pred=pm, var_state_pred=pS, x_meas=g["xm"], mean_meas=g["d"], wgt_meas=g["Wm"],
                            var_meas=g["Vm"])
        fm, fS = ktv.forecast(mean_state_pred=pm, var_state_pred=pS, mean_meas=g["d"], wgt_meas=g["Wm"],
                              var_meas=g["Vm"])
        sm, sS = ktv.smooth_mv(mean_state_next=g["mun"], var_state_next=g["Sn"], mean_state_filt=um,
                               var_state_filt=uS, mean_state_pred=pm, var_state_pred=pS, wgt_state=g["Qm"])
        ssm, ssS = ktv.smooth_sim(x_state_next=g["xn"], mean_state_filt=um, var_state_filt=uS, mean_state_pred=pm,
                                  var_state_pred=pS, wgt_state=g["Qm"])
        cA, cb, cV = ktv.smooth_cond(mean_state_filt=um, var_state_filt=uS, mean_state_pred=pm, var_state_pred=pS,
                                     wgt_state=g["Qm"])
        lp = rb.utils.multivariate_normal_logpdf(g["xm"], fm, fS)
        for n, val in zip(got, (pm, pS, um, uS, fm, fS, sm, sS, ssm, ssS, cA, cb, cV, lp)):
            got[n].append(_np(val) if hasattr(val, "detach") else np.asarray(val))
    for n, val in got.items():
        assert P.maxnorm_rel(np.stack(val), G["kal_" + n]) < TOL, n
    lp = [float(rb.utils.multivariate_normal_logpdf(G["lpcut_in_x"][i], np.zeros(2), G["lpcut_in_cov"][i]))
          for i in range(5)]
    assert np.allclose(lp, G["lpcut"], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_cuda_sampling_factor_squares_to_the_smoothing_covariance(rb):
    """The reference draws with the SVD factor U sqrt(s) of the (singular) smoothing covariance (src/rodeo/solve.py:179,
    182-186), whose column signs are LAPACK's: its draw vector `fitz_sim_x` cannot be reproduced bit for bit by any other
    factor, only in distribution.  What makes the CUDA draws equal in distribution is A A^T = V for the factor the
    kernels use -- checked here on every smoothing covariance of the `fitz_sim_*` golden set-up (filtered moments of the
    golden inputs, conditioned on the reference's own draws), to 1e-10 of max|V|; and the conditional means / covariances
    the CUDA primitives produce on the way equal the oracle's to 1e-10."""
    import torch
    pr, _ = prob("fitz")
    model, t_max, N = grid("fitz")
    mp, vp, mf, vf = orc.solve_filter(orc.MODELS[model], pr["W"], pr["X0"], 0.0, t_max, N, orc.interrogate_kramer,
                                      pr["Q"], pr["R"], pr["theta"])
    x = G["fitz_sim_x"]                                                    # the reference's draws (B, N+1, nb, p)
    B, _, nb, p = x.shape
    Q = np.broadcast_to(pr["Q"], (B, N - 1, nb, p, p))
    # smooth_sim at t = 1 .. N-1 conditions on x[t+1] with filt[t], pred[t+1]   (solve.py:170-178)
    m, V = rb.kalmantv.standard.smooth_sim(x[:, 2:N + 1], mf[:, 1:N], vf[:, 1:N], mp[:, 2:N + 1], vp[:, 2:N + 1], Q)
    om, oV = orc.smooth_sim(x[:, 2:N + 1], mf[:, 1:N], vf[:, 1:N], mp[:, 2:N + 1], vp[:, 2:N + 1], Q)
    assert P.maxnorm_rel(_np(m), om) < TOL and P.maxnorm_rel(_np(V), oV) < TOL
    Vall = torch.cat([V.reshape(-1, p, p), torch.as_tensor(vf[:, N].reshape(-1, p, p), device=V.device)])
    A = rb.kalmantv.standard.psd_factor(Vall)
    AAt = _np(A @ A.transpose(-1, -2))
    Vn = _np(Vall)
    scale = np.abs(Vn).max(axis=(1, 2), keepdims=True)
    err = float(np.max(np.abs(AAt - Vn) / scale))
    print(f"sampling factor: max |A A^T - V| / max|V| = {err:.2e} over {len(Vn)} covariances")
    assert np.array_equal(np.triu(_np(A), 1), np.zeros_like(Vn)) and err < TOL
