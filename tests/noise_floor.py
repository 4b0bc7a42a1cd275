"""Per-theta parity gate for the dalton log-likelihood (test infrastructure).

dalton is a difference of two sums of ~N terms z^2/S + log S whose residuals z ~ sqrt(S) ~ 1e-3 are differences of
O(1) quantities fed back through gains ~1/S: a faithful float64 evaluation of the reference recursion is only
reproducible to a theta-dependent noise floor (median ~1e-12, a heavy tail of ill-conditioned thetas up to ~1e-8 at
N = 800).  The floor is MEASURED per theta: `exact` is the same C restatement compiled in x87 long double
(oracle/c_port.dalton_ld, 64-bit mantissa), `oracle` a float64 evaluation of the reference algorithm.  The gate, per
theta, with err(x) = |x - exact| / max(1, |exact|):

        err(kernel)  <=  2 * err(oracle) + 1e-10

i.e. 1e-10 (BASELINE north_star) where the quantity is well conditioned, and no more than twice the reference
arithmetic's own float64 error where it is not.
"""
import json
import os

import numpy as np

TOL = 1e-10


def rel(x, exact):
    return np.abs(np.asarray(x, dtype=np.float64) - exact) / np.maximum(1.0, np.abs(exact))


def gate(kernel, oracle, exact, tol=TOL, factor=2.0):
    """stats of the per-theta gate; `ok` iff no theta violates it"""
    ek, eo = rel(kernel, exact), rel(oracle, exact)
    bound = factor * eo + tol
    bad = ek > bound
    q = lambda e: {k: float(np.quantile(e, v)) for k, v in (("median", 0.5), ("p99", 0.99), ("p999", 0.999), ("max", 1.0))}
    return {
        "n": int(ek.size), "gate": f"|kernel-exact| <= {factor:g}*|oracle-exact| + {tol:g} per theta (rel. to max(1,|exact|))",
        "n_violations": int(bad.sum()), "worst_excess": float(np.max(ek - bound)),
        "kernel_vs_exact": q(ek), "oracle_vs_exact": q(eo), "kernel_vs_oracle": q(rel(kernel, np.asarray(oracle))),
        "n_kernel_over_1e-10": int((ek > tol).sum()), "n_oracle_over_1e-10": int((eo > tol).sum()),
        "ok": bool(not bad.any()),
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
