"""Parity gate for the dalton log-likelihood against its measured float64 noise floor (test infrastructure).

dalton is a difference of two sums of ~N terms z^2/S + log S whose residuals z ~ sqrt(S) ~ 1e-3 are differences of
O(1) quantities fed back through gains ~1/S: a faithful float64 evaluation of the reference recursion is only
reproducible to a theta-dependent noise floor (median ~1e-12, a heavy tail of ill-conditioned thetas up to ~1e-8 at
N = 800).  The floor is MEASURED per theta: `exact` is the same C restatement compiled in x87 long double
(oracle/c_port.dalton_ld, 64-bit mantissa), `oracle` a float64 evaluation of the reference algorithm, and
err(x) = |x - exact| / max(1, |exact|).

Two float64 evaluations with different operation orders draw INDEPENDENT rounding noise of the same theta-dependent
scale, so a per-theta bound  err(kernel) <= 2 err(oracle) + 1e-10   is violated by chance wherever the
floor exceeds 1e-10: P(|X| > 2|Y|) = (2/pi) atan(1/2) = 0.30 for iid centred normals.  Measured on the bench workload
(65,536 thetas): 566 violations = 0.86 % of the batch = 30 % of the 1,861 thetas whose kernel error exceeds 1e-10 --
exactly that chance level.  The per-theta count is therefore reported, and what is asserted is

  * distribution dominance: at the median, the 90 %, 99 % and 99.9 % quantiles (those the sample size resolves)
        Q(err(kernel)) <= 2 Q(err(oracle)) + 1e-10,   and   max err(kernel) <= 4 max err(oracle) + 1e-10;
  * the number of per-theta violations stays below 2 % of the batch + 3 (chance level ~0.9 % on the bench workload;
    the "+ 3" is for batches of ~100 thetas, where two or three ill-conditioned thetas are already 2-3 %).

Where the quantity is well conditioned (err(oracle) << 1e-10) this is the 1e-10 of BASELINE's north_star.
"""
import json
import os

import numpy as np

TOL = 1e-10


def rel(x, exact):
    return np.abs(np.asarray(x, dtype=np.float64) - exact) / np.maximum(1.0, np.abs(exact))


def gate(kernel, oracle, exact, tol=TOL, factor=2.0):
    """statistics of the gate described in the module docstring; `ok` iff it holds"""
    ek, eo = rel(kernel, exact), rel(oracle, exact)
    n = int(ek.size)
    bad = ek > factor * eo + tol
    qs = [("median", 0.5), ("p90", 0.9)] + ([("p99", 0.99)] if n >= 1000 else []) + ([("p999", 0.999)] if n >= 10000 else [])
    q = lambda e: {**{k: float(np.quantile(e, v)) for k, v in qs}, "max": float(e.max())}
    qk, qo = q(ek), q(eo)
    dominated = all(qk[k] <= factor * qo[k] + tol for k, _ in qs) and qk["max"] <= 2 * factor * qo["max"] + tol
    frac = float(bad.mean())
    return {
        "n": n,
        "gate": f"quantiles(|kernel-exact|) <= {factor:g}*quantiles(|oracle-exact|) + {tol:g}, max <= {2 * factor:g}x; "
                f"per-theta violations of |kernel-exact| <= {factor:g}*|oracle-exact| + {tol:g} at most 2% of n + 3 "
                f"(errors relative to max(1,|exact|))",
        "kernel_vs_exact": qk, "oracle_vs_exact": qo, "kernel_vs_oracle": q(rel(kernel, np.asarray(oracle))),
        "n_violations_per_theta": int(bad.sum()), "violation_frac": frac,
        "worst_excess": float(np.max(ek - (factor * eo + tol))),
        "n_kernel_over_1e-10": int((ek > tol).sum()), "n_oracle_over_1e-10": int((eo > tol).sum()),
        "ok": bool(dominated and bad.sum() <= 0.02 * n + 3 and np.isfinite(ek).all()),
    }


def report(name, st):
    print(f"{name}: {json.dumps(st)}")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        try:
            with open(os.path.join(out, "parity_stats.jsonl"), "a") as f:
                f.write(json.dumps({"name": name, **st}) + "\n")
        except OSError:
            pass
