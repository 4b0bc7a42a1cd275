#!/usr/bin/env python
"""Regenerate tests/golden/oracle_vectors.npz.

These vectors are produced by the NumPy ORACLE (oracle/rodeo_oracle.py), not by the reference: the reference is pure
JAX and cannot run in this image (no jax, no network), and it ships no fixtures of its own.  They freeze the oracle's
outputs on small seeded problems so that (a) an accidental change of the oracle is caught on the CPU, and (b) the GPU
path is also compared against numbers that were committed, not recomputed in the same process.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import problems as P  # noqa: E402
from oracle import rodeo_oracle as orc  # noqa: E402


def build():
    out = {}
    pr = P.fitz_problem(4, n_steps=60, t_max=3.0, seed=123)
    ob = P.fitz_obs(pr, None, n_obs=4)
    mdl = orc.MODELS["fitzhugh_nagumo"]
    args = (mdl, pr["W"], pr["X0"], 0.0, 3.0, 60, orc.interrogate_kramer, (pr["Q"], pr["R"]), pr["theta"])
    obs = (ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    out["fitz_mean"], out["fitz_var"] = orc.solve_mv(*args)
    out["fitz_dalton"] = orc.dalton(*args, *obs)
    out["fitz_fenrir"] = orc.fenrir(*args, *obs)
    out["fitz_dalton_mean"], out["fitz_dalton_var"] = orc.dalton_solve_mv(*args, *obs)
    out["fitz_fenrir_mean"], out["fitz_fenrir_var"] = orc.fenrir_solve_mv(*args, *obs)
    zs = np.random.default_rng(5).standard_normal((4, 61, 2, 3))
    out["fitz_sim_z"] = zs
    out["fitz_sim"] = orc.solve_sim(*args, z_smooth=zs, factor="ldl")
    pr2 = P.second_order_problem(3, n_steps=80, t_max=4.0, sigma=0.1, seed=123)
    ob2 = P.second_order_obs(pr2, n_obs=5)
    args2 = (orc.MODELS["second_order_sin"], pr2["W"], pr2["X0"], 0.0, 4.0, 80, orc.interrogate_kramer,
             (pr2["Q"], pr2["R"]), pr2["theta"])
    out["so_mean"], out["so_var"] = orc.solve_mv(*args2)
    out["so_fenrir"] = orc.fenrir(*args2, ob2["obs_data"], ob2["obs_times"], ob2["obs_weight"], ob2["obs_var"])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **build())
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"))
