"""Pins the C/OpenMP oracle port (cpu_baseline) to the canonical NumPy oracle.  CPU only."""
import time

import numpy as np
import pytest

import problems as P
from oracle import c_port, rodeo_oracle as orc

ORC_INTERR = {"kramer": orc.interrogate_kramer, "schober": orc.interrogate_schober, "rodeo": orc.interrogate_rodeo}


def _ll_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


@pytest.mark.parametrize("interr", ["kramer", "rodeo"])
def test_c_dalton_matches_numpy_oracle(interr):
    pr = P.fitz_problem(24, n_steps=200, t_max=10.0, seed=2)
    ob = P.fitz_obs(pr, None, n_obs=11)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 10.0, 200, ORC_INTERR[interr],
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"],
                      ob["obs_var"])
    ind = orc.obs_index(0.0, 10.0, 200, ob["obs_times"])
    got = c_port.dalton("fitzhugh_nagumo", interr, pr["W"], pr["X0"], 0.0, 10.0, 200, pr["Q"], pr["R"],
                        pr["theta"], ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"])
    assert _ll_err(got, want) < 2e-9       # float64 noise floor of the log-likelihood, see test_gpu_parity.py


@pytest.mark.parametrize("name,setup", [("fitzhugh_nagumo", lambda: P.fitz_problem(8, 150, 7.5, seed=3)),
                                        ("lorenz63", lambda: P.lorenz_problem(4, 200, 1.0, seed=3)),
                                        ("second_order_sin", lambda: P.second_order_problem(6, 300, 10.0, 0.01, seed=3))])
def test_c_solve_mv_matches_numpy_oracle(name, setup):
    pr = setup()
    N, tm = pr["n_steps"], pr["t_max"]
    om, ov = orc.solve_mv(orc.MODELS[name], pr["W"], pr["X0"], 0.0, tm, N, orc.interrogate_kramer,
                          (pr["Q"], pr["R"]), pr["theta"])
    m, v = c_port.solve_mv(name, "kramer", pr["W"], pr["X0"], 0.0, tm, N, pr["Q"], pr["R"], pr["theta"])
    assert P.maxnorm_rel(m, om) < 1e-10 and P.maxnorm_rel(v, ov) < 1e-10


def test_c_port_is_thread_count_invariant():
    pr = P.fitz_problem(16, n_steps=100, t_max=5.0, seed=4)
    ob = P.fitz_obs(pr, None, n_obs=6)
    ind = orc.obs_index(0.0, 5.0, 100, ob["obs_times"])
    a = c_port.dalton("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"], 0.0, 5.0, 100, pr["Q"], pr["R"], pr["theta"],
                      ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"], n_threads=1)
    b = c_port.dalton("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"], 0.0, 5.0, 100, pr["Q"], pr["R"], pr["theta"],
                      ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"], n_threads=0)
    assert np.array_equal(a, b)
