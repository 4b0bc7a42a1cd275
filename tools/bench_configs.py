#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json config on one B200 (the driver's bench.py covers configs[1]).

Prints one JSON line per config: theta*steps/s, ms per call, and the algorithmic roofline fraction from the dense
flop / byte counts of SURVEY.md section 8(d).  Inputs are synthetic (tests/problems.py).
"""
import argparse
import functools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P  # noqa: E402
import rodeo_b200 as rb  # noqa: E402
from rodeo_b200 import _lib  # noqa: E402
import ctypes  # noqa: E402


def timeit(fn, reps, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del out
    return float(np.median(ts))


def cold_ms(fn, lib):
    """one call with the covariance-schedule cache emptied first: what the first call of a (prior, n_steps) pays"""
    lib.rodeo_b200_schedule_clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record()
    torch.cuda.synchronize()
    del out
    return float(e0.elapsed_time(e1))


def run(only="", reps=5, scale=1.0, quiet=False):
    """time the configs named in `only` (comma separated; empty = all float64 ones); returns the list of records"""
    class A:
        pass
    args = A()
    args.only, args.reps, args.scale = only, reps, scale
    dev = torch.device("cuda", torch.cuda.current_device())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    peak = ctypes.c_double(0.0)
    lib.rodeo_b200_fp64_peak_probe(5, ctypes.byref(peak))
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    kr = rb.interrogate.interrogate_kramer
    chk = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="standard")
    D = lambda a: torch.as_tensor(a, device=dev)
    res = []

    def report(name, B, N, ms, flops, bytes_, extra=None):
        rate = B * N / (ms * 1e-3)
        t_f = flops * B * N / (peak.value * 1e12); t_b = bytes_ * B * N / (hbm * 1e9)
        bound = "fp64" if t_f >= t_b else "hbm"
        line = {"config": name, "B": B, "n_steps": N, "ms": ms, "theta_steps_per_s": rate,
                "alg_flops_per_theta_step": flops, "alg_bytes_per_theta_step": bytes_, "bound": bound,
                "roofline_frac": max(t_f, t_b) / (ms * 1e-3), "fp64_peak_tflops": peak.value, "hbm_gbs": hbm}
        if extra:
            line.update(extra)
        if "moved_bytes_per_theta_step" in line:
            # bytes the kernel moves by design (compulsory output + the history its backward sweep needs, each way):
            # the honest HBM view for kernels that no longer execute the dense flop count
            line["hbm_frac_design_traffic"] = line["moved_bytes_per_theta_step"] * B * N / (ms * 1e-3) / (hbm * 1e9)
        if not quiet:
            print(json.dumps(line), flush=True)
        res.append(line)

    # solve_sim under chkrebtii runs over a cached covariance schedule unless RODEO_SIM_SCHEDULE=0 (rodeo_sched.cuh): the
    # kernel then carries block means only, so SURVEY 8(d)'s dense flop count is no longer what is executed --
    # roofline_frac (dense count / measured DFMA peak) is kept for comparison, hbm_frac_design_traffic is the bound
    sched = os.environ.get("RODEO_SIM_SCHEDULE", "1") != "0"
    kname = ("solve_sim_sched_kernel (block means over a cached covariance schedule)" if sched
             else "solve_sim_kernel / solve_sim_bl_kernel (full covariance recursion per theta)")
    want = lambda k: (k in args.only.split(",")) if args.only else not (k.endswith("f32") or k == "C5x")
    sc = args.scale

    if want("C1"):
        B = int(65536 * sc); pr = P.fitz_problem(B, seed=0)
        X0, th = D(pr["X0"]), D(pr["theta"])
        f = lambda: rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0, 800, kr,
                                prior_pars=(pr["Q"], pr["R"]), theta=th)
        report("C1 FN solve_mv kramer", B, 800, timeit(f, args.reps, flush), 1013.0, 192.0)
        del X0, th; torch.cuda.empty_cache()
    if want("C2"):
        B = int(65536 * sc); pr = P.fitz_problem(B, seed=0); ob = P.fitz_obs(pr, None)
        X0, th = D(pr["X0"]), D(pr["theta"])
        obd = {k: D(v) if k != "obs_times" else v for k, v in ob.items()}
        f = lambda: rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0, 800, kr,
                                        prior_pars=(pr["Q"], pr["R"]), theta=th, **obd)
        report("C2 FN dalton kramer", B, 800, timeit(f, args.reps, flush), 963.0, 0.0)
    if want("C1f32"):
        B = int(65536 * sc); pr = P.fitz_problem(B, seed=0)
        X0, th = D(pr["X0"].astype(np.float32)), D(pr["theta"].astype(np.float32))
        f = lambda: rb.solve_mv(None, rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0, 800, kr,
                                prior_pars=(pr["Q"], pr["R"]), theta=th)
        report("C1f32 FN solve_mv kramer, float32 storage / mixed arithmetic", B, 800, timeit(f, args.reps, flush),
               1013.0, 96.0)
        del X0, th; torch.cuda.empty_cache()
    if want("C2f32"):
        B = int(65536 * sc); pr = P.fitz_problem(B, seed=0); ob = P.fitz_obs(pr, None)
        X0, th = D(pr["X0"].astype(np.float32)), D(pr["theta"].astype(np.float32))
        obd = {k: D(v) if k != "obs_times" else v for k, v in ob.items()}
        f = lambda: rb.inference.dalton(None, rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0, 800, kr,
                                        prior_pars=(pr["Q"], pr["R"]), theta=th, **obd)
        report("C2f32 FN dalton kramer, float32 storage / mixed arithmetic", B, 800, timeit(f, args.reps, flush),
               963.0, 0.0)
    if want("C3"):
        B = int(65536 * sc); pr = P.lorenz_problem(B, seed=0)
        X0, th = D(pr["X0"]), D(pr["theta"])
        f = lambda: rb.solve_sim(np.array([1, 2], dtype=np.uint32), rb.models.lorenz63, pr["W"], X0, 0.0, 20.0, 4000,
                                 chk, prior_pars=(pr["Q"], pr["R"]), theta=th)
        ms = timeit(f, max(2, args.reps // 2), flush)
        x = f(); fin = bool(torch.isfinite(x).all().item()); del x
        report("C3 Lorenz63 solve_sim chkrebtii (4,096 theta x 16 draws)", B, 4000, ms, 1539.0, 72.0,
               {"finite": fin, "moved_bytes_per_theta_step": 72.0 + (2 * 72.0 if sched else 2 * 216.0 / 2), "kernel": kname,
                "first_call_ms_schedule_build_included": cold_ms(f, lib) if sched else None})
        del X0, th; torch.cuda.empty_cache()
    if want("C4"):
        B = int(16384 * sc); pr = P.second_order_problem(B, seed=0); ob = P.second_order_obs(pr)
        X0, th = D(pr["X0"]), D(pr["theta"])
        obd = {k: D(v) if k != "obs_times" else v for k, v in ob.items()}
        f = lambda: rb.inference.fenrir(None, rb.models.second_order_sin, pr["W"], X0, 0.0, 10.0, 2000, kr,
                                        prior_pars=(pr["Q"], pr["R"]), theta=th, **obd)
        report("C4 second-order fenrir kramer", B, 2000, timeit(f, args.reps, flush), 1230.0, 0.0)
    if want("C5"):
        # BASELINE configs[4]: one pseudo-marginal iteration = solve_sim + Gaussian observation log-likelihood per
        # particle; the fused kernel never writes the trajectories (SURVEY 8(d) C5 "~0 B, log-lik only")
        B = int(32768 * sc); pr = P.fitz_problem(B, seed=0); ob = P.fitz_obs(pr, None)
        X0, th = D(pr["X0"]), D(pr["theta"])
        Y = D(ob["obs_data"][:, :, 0])
        f = lambda: rb.solve_sim_loglik(np.array([5, 6], dtype=np.uint32), rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0,
                                        40.0, 800, chk, prior_pars=(pr["Q"], pr["R"]), theta=th, obs_data=Y,
                                        obs_times=ob["obs_times"], noise_sd=0.0707)
        report("C5 FN solve_sim chkrebtii + obs log-lik, fused, no Xt (one GPU's 32,768 of 262,144 particles)", B, 800,
               timeit(f, args.reps, flush), 1035.0, 0.0,
               {"moved_bytes_per_theta_step": 2 * 48.0 if sched else 2 * 144.0, "kernel": kname,
                "first_call_ms_schedule_build_included": cold_ms(f, lib) if sched else None})
    if want("C5x"):
        B = int(32768 * sc); pr = P.fitz_problem(B, seed=0)
        X0, th = D(pr["X0"]), D(pr["theta"])
        f = lambda: rb.solve_sim(np.array([5, 6], dtype=np.uint32), rb.models.fitzhugh_nagumo, pr["W"], X0, 0.0, 40.0,
                                 800, chk, prior_pars=(pr["Q"], pr["R"]), theta=th)
        report("C5x FN solve_sim chkrebtii writing Xt (32,768 particles)", B, 800,
               timeit(f, args.reps, flush), 1035.0, 48.0)
    del flush
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (debugging)")
    args = ap.parse_args()
    res = run(args.only, args.reps, args.scale)
    out = os.path.join(ROOT, "gpurun_out", "bench_configs.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
