#!/usr/bin/env python
"""Experiment: end-to-end time of rodeo_b200_dalton_f64_host (pinned host buffers) for several chunk splits
(RODEO_HOST_CHUNKS / RODEO_HOST_SPLIT).  usage: python tools/e2e_split.py"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P
import rodeo_b200 as rb
from rodeo_b200 import _host, _lib
B, N = 65536, 800
pr = P.fitz_problem(B, seed=0); ob = P.fitz_obs(pr, None)
lib = _lib.load()
pb = _host.Problem(None, rb.models.fitzhugh_nagumo, pr["W"], pr["X0"], 0.0, 40.0, N, rb.interrogate.interrogate_kramer,
                   (pr["Q"], pr["R"]), None, None, "standard", {"theta": pr["theta"]}, host_inputs=True)
pb.set_obs(ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
h_x0 = torch.from_numpy(pr["X0"]).pin_memory(); h_th = torch.from_numpy(pr["theta"]).pin_memory()
h_out = torch.empty((B,), dtype=torch.float64).pin_memory()
def step():
    rc = lib.rodeo_b200_dalton_f64_host(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                        ctypes.c_void_p(h_x0.data_ptr()), ctypes.c_void_p(h_th.data_ptr()),
                                        _host.ptr(pb.obs_ind), _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight),
                                        _host.ptr(pb.obs_var), ctypes.c_void_p(h_out.data_ptr()))
    _lib.check(rc, "dalton_host")
ref = None
for name, env in [("chunks=1", {"RODEO_HOST_CHUNKS": "1"}), ("chunks=2", {"RODEO_HOST_CHUNKS": "2"}),
                  ("split .125,.25,.625", {"RODEO_HOST_SPLIT": "0.125,0.25,0.625"}),
                  ("split .0625,.1875,.75", {"RODEO_HOST_SPLIT": "0.0625,0.1875,0.75"}),
                  ("split .125,.375,.5", {"RODEO_HOST_SPLIT": "0.125,0.375,0.5"}),
                  ("split .25,.75", {"RODEO_HOST_SPLIT": "0.25,0.75"}),
                  ("split .125,.875", {"RODEO_HOST_SPLIT": "0.125,0.875"}),
                  ("split .0625,.125,.25,.5625", {"RODEO_HOST_SPLIT": "0.0625,0.125,0.25,0.5625"}),
                  ("chunks=2 again", {"RODEO_HOST_CHUNKS": "2"})]:
    for k in ("RODEO_HOST_CHUNKS", "RODEO_HOST_SPLIT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(5): step()
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(40): step()
        ts.append((time.perf_counter() - t0) / 40)
    if ref is None: ref = h_out.numpy().copy()
    print(f"{name:28s} {min(ts)*1e3:.4f} ms  {B*N/min(ts)/1e9:.1f} G  same={np.array_equal(ref, h_out.numpy())}", flush=True)
