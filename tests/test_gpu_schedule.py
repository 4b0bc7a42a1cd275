"""GPU: solve_sim over a covariance schedule (rodeo_b200/csrc/rodeo_sched.cuh) against the full kernels and the oracle.

Under interrogate_chkrebtii / _schober / _rodeo the reference's covariance recursion does not see the state
(src/rodeo/interrogate.py:13-60, 87-115; src/rodeo/solve.py:59-88), so the library tabulates it once and the per-theta
kernel carries the block means only.  The table holds bitwise what every theta of the full kernels computes for itself,
and the mean arithmetic keeps their operation order: without a per-theta prior scale the draws must be BITWISE those of
solve_sim_kernel.  RODEO_SIM_SCHEDULE=0 selects the full kernels.
"""
import functools

import numpy as np
import pytest

import problems as P
from oracle import rodeo_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rodeo_b200
    from rodeo_b200 import _lib
    _lib.load()
    return rodeo_b200


def _np(t):
    return t.detach().cpu().numpy()


def _interr(rb, name):
    f = getattr(rb.interrogate, "interrogate_" + name)
    return functools.partial(f, kalman_type="standard") if name == "chkrebtii" else f


def _problem(name, B):
    if name == "fitzhugh_nagumo":
        return P.fitz_problem(B, n_steps=57, t_max=2.85, seed=71)
    if name == "lorenz63":
        return P.lorenz_problem(B, n_steps=40, t_max=0.2, sigma=1.0, seed=72)
    return P.second_order_problem(B, n_steps=64, t_max=1.0, sigma=0.1, seed=73)


def _both(monkeypatch, fn, lanes="0"):
    """fn() through the full kernels and through the schedule, both with one lane per theta (lanes = "0") or one lane
    per (theta, block) where the model allows it (lanes = "1")"""
    monkeypatch.setenv("RODEO_SIM_BLOCK_LANES", lanes)
    monkeypatch.setenv("RODEO_SIM_SCHEDULE", "0")
    full = fn()
    monkeypatch.setenv("RODEO_SIM_SCHEDULE", "1")
    sched = fn()
    return full, sched


@pytest.mark.parametrize("lanes", ["0", "1"])
@pytest.mark.parametrize("interr", ["chkrebtii", "schober", "rodeo"])
@pytest.mark.parametrize("model,B", [("fitzhugh_nagumo", 40), ("lorenz63", 23), ("second_order_sin", 33),
                                     ("fitzhugh_nagumo", 1)])
def test_schedule_draws_are_bitwise_the_full_kernels(rb, monkeypatch, model, B, interr, lanes):
    pr = _problem(model, B)
    key = np.array([11, 5], dtype=np.uint32)
    run = lambda: _np(rb.solve_sim(key, getattr(rb.models, model), pr["W"], pr["X0"], 0.0, pr["t_max"], pr["n_steps"],
                                   _interr(rb, interr), prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"]))
    full, sched = _both(monkeypatch, run, lanes)
    assert np.isfinite(full).all() and np.array_equal(sched[:, 0], pr["X0"])
    assert np.array_equal(full, sched)
    if lanes == "1":                                   # ... and the lane mapping does not change a single bit either
        monkeypatch.setenv("RODEO_SIM_BLOCK_LANES", "0")
        assert np.array_equal(run(), sched)


def test_schedule_injected_normals_bitwise_and_against_the_oracle(rb, monkeypatch):
    B, N = 24, 120
    pr = P.fitz_problem(B, n_steps=N, t_max=6.0, seed=74)
    rng = np.random.default_rng(75)
    zs, zi = rng.standard_normal((B, N + 1, 2, 3)), rng.standard_normal((B, N, 1, 2, 3))
    run = lambda: _np(rb.solve_sim(0, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 6.0, N,
                                   _interr(rb, "chkrebtii"), prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"],
                                   _z_smooth=zs, _z_interr=zi))
    full, sched = _both(monkeypatch, run)
    assert np.array_equal(full, sched)
    want = orc.solve_sim(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 6.0, N,
                         functools.partial(orc.interrogate_chkrebtii, factor="ldl"), (pr["Q"], pr["R"]), pr["theta"],
                         z_smooth=zs, z_interrogate=zi[:, :, 0], factor="ldl")
    e = P.maxnorm_rel(sched, want)
    print(f"schedule solve_sim vs oracle, injected normals: {e:.2e}")
    assert e < 1e-8


def test_schedule_dense_instantiation(rb, monkeypatch):
    """a prior weight that is not unit upper triangular and a general ode_weight run the dense instantiation"""
    B, N = 19, 48
    pr = P.fitz_problem(B, n_steps=N, t_max=2.4, seed=76)
    Q = pr["Q"].copy(); Q[:, 0, 0] = 0.98; Q[:, 2, 1] = 0.01
    W = pr["W"].copy(); W[:, 0, 0] = 0.05
    for interr in ("chkrebtii", "schober", "rodeo"):
        run = lambda: _np(rb.solve_sim(np.array([1, 2], dtype=np.uint32), rb.models.fitzhugh_nagumo, W, pr["X0"], 0.0,
                                       2.4, N, _interr(rb, interr), prior_pars=(Q, pr["R"]), theta=pr["theta"]))
        full, sched = _both(monkeypatch, run)
        assert np.isfinite(full).all() and np.array_equal(full, sched), interr


def test_schedule_with_a_per_theta_prior_scale(rb, monkeypatch):
    """sigma as part of theta (docs/examples/parameter.md:218-236, the reference's pseudo-marginal walk-through): every
    covariance of block b scales by sigma_b^2 exactly, so the tabulated unit-scale factors are scaled per theta."""
    B, N = 32, 100
    pr = P.fitz_problem(B, n_steps=N, t_max=5.0, seed=77)
    sig = 0.1 * np.exp(0.4 * np.random.default_rng(78).standard_normal((B, 2)))
    Q, Rb = rb.prior.ibm_init(5.0 / N, 3, sig)
    rng = np.random.default_rng(79)
    zs, zi = rng.standard_normal((B, N + 1, 2, 3)), rng.standard_normal((B, N, 1, 2, 3))
    for interr in ("chkrebtii", "rodeo"):
        run = lambda: _np(rb.solve_sim(0, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, N,
                                       _interr(rb, interr), prior_pars=(Q, Rb), theta=pr["theta"], _z_smooth=zs,
                                       _z_interr=zi))
        full, sched = _both(monkeypatch, run)
        e = P.maxnorm_rel(sched, full)
        print(f"per-theta prior scale, {interr}: schedule vs full kernel {e:.2e}")
        assert e < 1e-12
    Ro = np.stack([orc.ibm_init(5.0 / N, 3, sig[k])[1] for k in range(B)])
    want = orc.solve_sim(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 5.0, N,
                         functools.partial(orc.interrogate_chkrebtii, factor="ldl"), (Q, Ro), pr["theta"],
                         z_smooth=zs, z_interrogate=zi[:, :, 0], factor="ldl")
    monkeypatch.setenv("RODEO_SIM_SCHEDULE", "1")
    x = _np(rb.solve_sim(0, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 5.0, N, _interr(rb, "chkrebtii"),
                         prior_pars=(Q, Rb), theta=pr["theta"], _z_smooth=zs, _z_interr=zi))
    assert P.maxnorm_rel(x, want) < 1e-8


@pytest.mark.parametrize("lanes", ["0", "1"])
def test_schedule_fused_loglik(rb, monkeypatch, lanes):
    N, tm, B = 160, 8.0, 77
    pr = P.fitz_problem(B, n_steps=N, t_max=tm, seed=37)
    ob = P.fitz_obs(pr, None, n_obs=9)
    Y = ob["obs_data"][:, :, 0]
    key = np.array([3, 9], dtype=np.uint32)
    args = (key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, tm, N, _interr(rb, "chkrebtii"))
    kw = dict(prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"], obs_data=Y, obs_times=ob["obs_times"], noise_sd=0.07)
    full, sched = _both(monkeypatch, lambda: _np(rb.solve_sim_loglik(*args, **kw)), lanes)
    assert np.array_equal(full, sched)
    ll2, x2 = rb.solve_sim_loglik(*args, return_draws=True, **kw)
    assert np.array_equal(_np(ll2), sched)
    ind = orc.obs_index(0.0, tm, N, ob["obs_times"])
    assert np.max(np.abs(orc.gauss_obs_loglik(_np(x2), ind, Y, 0.07) - sched) / np.maximum(1, np.abs(sched))) < 1e-12


def test_schedule_is_built_once_per_prior_and_grid(rb, monkeypatch):
    from rodeo_b200 import _lib
    lib = _lib.load()
    monkeypatch.setenv("RODEO_SIM_SCHEDULE", "1")
    lib.rodeo_b200_schedule_clear()
    n0 = lib.rodeo_b200_schedule_builds()
    pr = P.fitz_problem(9, n_steps=30, t_max=1.5, seed=80)
    chk = _interr(rb, "chkrebtii")
    call = lambda key, N, R: _np(rb.solve_sim(np.array([key, 0], dtype=np.uint32), rb.models.fitzhugh_nagumo, pr["W"],
                                              pr["X0"], 0.0, 1.5, N, chk, prior_pars=(pr["Q"], R), theta=pr["theta"]))
    a = call(1, 30, pr["R"])
    assert lib.rodeo_b200_schedule_builds() == n0 + 1
    b = call(2, 30, pr["R"])                       # other key, same prior and grid: table re-used
    assert lib.rodeo_b200_schedule_builds() == n0 + 1 and not np.array_equal(a, b)
    assert np.array_equal(call(1, 30, pr["R"]), a)
    call(1, 29, pr["R"])                           # other grid
    call(1, 30, 1.5 * pr["R"])                     # other prior
    assert lib.rodeo_b200_schedule_builds() == n0 + 3
    lib.rodeo_b200_schedule_clear()
    assert np.array_equal(call(1, 30, pr["R"]), a)
    assert lib.rodeo_b200_schedule_builds() == n0 + 4
    # more distinct schedules than the cache holds: eviction keeps results intact
    outs = [call(1, 10 + k, pr["R"]) for k in range(20)]
    assert all(np.array_equal(call(1, 10 + k, pr["R"]), outs[k]) for k in (0, 7, 19))


def test_schedule_float32(rb, monkeypatch):
    B, N = 21, 80
    pr = P.fitz_problem(B, n_steps=N, t_max=4.0, seed=81)
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    run = lambda: _np(rb.solve_sim(np.array([4, 4], dtype=np.uint32), rb.models.fitzhugh_nagumo, pr["W"],
                                   f32(pr["X0"]), 0.0, 4.0, N, _interr(rb, "chkrebtii"), prior_pars=(pr["Q"], pr["R"]),
                                   theta=f32(pr["theta"])))
    full, sched = _both(monkeypatch, run)
    e = P.maxnorm_rel(sched, full)
    print(f"float32 schedule vs full kernel: {e:.2e}")
    assert full.dtype == np.float32 and e < 5e-6


@pytest.mark.parametrize("lanes", ["0", "1"])
@pytest.mark.parametrize("B", [1, 5, 33, 47])
def test_schedule_ragged_batches_and_output_canaries(rb, monkeypatch, B, lanes):
    """Ragged batch sizes through both lane mappings of the schedule kernels, straight through the C ABI: rows equal those
    of a larger batch, the draws are fully written, nothing is written outside the output / workspace buffers (guard
    words on both sides; compute-sanitizer is closed on this pool)."""
    import ctypes
    import torch
    from rodeo_b200 import _host, _lib
    monkeypatch.setenv("RODEO_SIM_SCHEDULE", "1")
    monkeypatch.setenv("RODEO_SIM_BLOCK_LANES", lanes)
    N, tm = 37, 1.85
    big = P.fitz_problem(64, n_steps=N, t_max=tm, seed=91)
    key = np.array([21, 4], dtype=np.uint32)
    chk = _interr(rb, "chkrebtii")
    args = (rb.models.fitzhugh_nagumo, big["W"])
    ref = rb.solve_sim(key, *args, big["X0"], 0.0, tm, N, chk, prior_pars=(big["Q"], big["R"]), theta=big["theta"])
    pb = _host.Problem(key, *args, big["X0"][:B], 0.0, tm, N, chk, (big["Q"], big["R"]), None, None, "standard",
                       {"theta": big["theta"][:B]})
    G, SENT = 4096, -777.25

    def guarded(n):
        t = torch.full((n + 2 * G,), SENT, dtype=torch.float64, device="cuda")
        return t, t[G:G + n]

    nx = B * (N + 1) * 6
    wsb = pb.lib.rodeo_b200_workspace_bytes(_lib.OP_SOLVE_SIM, ctypes.byref(pb.c), 8)
    (gx, x), (gw, w) = guarded(nx), guarded(max(wsb // 8, 1))
    rc = pb.fn("solve_sim")(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R), _host.ptr(pb.x0),
                            _host.ptr(pb.theta), None, None, _host.ptr(x), _host.ptr(w), wsb, pb.stream())
    _lib.check(rc, "solve_sim")
    torch.cuda.synchronize()
    for g, n in ((gx, nx), (gw, max(wsb // 8, 1))):
        assert bool((g[:G] == SENT).all()) and bool((g[G + n:] == SENT).all()), "write outside the buffer"
    assert not bool((x == SENT).any()), "draws not fully written"
    assert torch.equal(x.view(B, N + 1, 2, 3), ref[:B])


@pytest.mark.parametrize("lanes", ["0", "1"])
@pytest.mark.parametrize("N", [1, 2, 3, 15, 16, 17, 32, 33])
def test_schedule_step_counts_at_the_chunk_edges(rb, monkeypatch, N, lanes):
    """Table rows are streamed through shared memory in chunks (16 backward rows, up to 32 forward rows): step counts
    below / at / just above a chunk, draws and fused log-likelihood, bitwise the full kernels."""
    B = 19
    pr = P.fitz_problem(B, n_steps=N, t_max=0.05 * N, seed=100 + N)
    key = np.array([7, N], dtype=np.uint32)
    obs_t = np.array([0.0, 0.05 * N])
    Y = np.array([[-1.0, 1.0], [-0.9, 0.9]])
    args = (key, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 0.05 * N, N, _interr(rb, "chkrebtii"))
    kw = dict(prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"])

    def run():
        ll, x = rb.solve_sim_loglik(*args, obs_data=Y, obs_times=obs_t, noise_sd=0.1, return_draws=True, **kw)
        return np.concatenate([_np(ll).ravel(), _np(x).ravel(), _np(rb.solve_sim(*args, **kw)).ravel(),
                               _np(rb.solve_sim_loglik(*args, obs_data=Y, obs_times=obs_t, noise_sd=0.1, **kw)).ravel()])
    full, sched = _both(monkeypatch, run, lanes)
    assert np.isfinite(full).all() and np.array_equal(full, sched)
