"""Experiment: solve_mv time with / without the output writes (mean_out / var_out = NULL skips the copy-out)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P
import rodeo_b200 as rb
from rodeo_b200 import _host, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pr = P.fitz_problem(B, seed=0)
pb = _host.Problem(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 40.0, 800, rb.interrogate.interrogate_kramer,
                   (pr["Q"], pr["R"]), None, None, "standard", {"theta": pr["theta"]})
dev = _host.device()
mean = torch.empty((B, 801, 2, 3), dtype=torch.float64, device=dev)
var = torch.empty((B, 801, 2, 3, 3), dtype=torch.float64, device=dev)
ws, n = pb.workspace(_lib.OP_SOLVE_MV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(m, v):
    rc = pb.lib.rodeo_b200_solve_mv_f64(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
        _host.ptr(pb.x0), _host.ptr(pb.theta), None, _host.ptr(m), _host.ptr(v), _host.ptr(ws), n, pb.stream())
    _lib.check(rc, "solve_mv")
for name, (m, v) in {"full": (mean, var), "mean only": (mean, None), "no output": (None, None)}.items():
    ts = []
    for i in range(6):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(m, v); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{name:10s} {np.median(ts[2:]):.3f} ms")
