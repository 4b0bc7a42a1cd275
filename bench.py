#!/usr/bin/env python
"""Benchmark of the rodeo filtering hot path on B200:  theta*steps / s  for the batched dalton log-likelihood.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): FitzHugh-Nagumo `rodeo.inference.dalton`, float64,
B = 65,536 thetas PER GPU (weak scaling; the theta batch shards with no data-path exchange), n_steps = 800 on
t in [0, 40], n_obs = 41, interrogate_kramer, IBM prior sigma = 0.1.  One "step" = one pass of the hot path over
the batch = one kernel launch (plus, for N > 1, the NCCL all-gather of the per-theta log-likelihoods).

One JSON line on stdout (rank 0).  `value` is device-resident throughput; `e2e` goes through the C ABI with HOST
buffers (rodeo_b200_dalton_f64_host: H2D of X0/theta/obs, kernel, D2H of the log-likelihoods, every step).
`--impl reference` times the CPU port of the reference algorithm (oracle/rodeo_oracle.c, all host threads) on a
bounded sample of the same workload: the reference's own JAX path cannot run in this image (no jax, no network).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems as P  # noqa: E402  (pure NumPy problem set-ups shared with the tests)

B_PER_GPU = 65536
N_STEPS = 800
N_OBS = 41
T_MAX = 40.0
METRIC = "theta_steps_per_sec"
UNIT = "theta*steps/s"
# dense algorithmic flops per theta*step of C2 (two filters, two blocks, p=3, m=1, obs every 20 steps)
FLOPS_PER_THETA_STEP = 963.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload(B, seed):
    pr = P.fitz_problem(B, n_steps=N_STEPS, t_max=T_MAX, sigma=0.1, seed=seed)
    return pr


def obs_for(pr, truth_mean):
    return P.fitz_obs(pr, truth_mean, n_obs=N_OBS, noise_var=0.005, seed=1)


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_problem(B):
    from oracle import c_port, rodeo_oracle as orc
    pr = workload(B, seed=0)
    pr0 = P.fitz_problem(1, N_STEPS, T_MAX, jitter=False)
    truth, _ = c_port.solve_mv("fitzhugh_nagumo", "kramer", pr0["W"], pr0["X0"], 0.0, T_MAX, N_STEPS, pr0["Q"],
                               pr0["R"], pr0["theta"])
    ob = obs_for(pr, truth[0])
    ind = orc.obs_index(0.0, T_MAX, N_STEPS, ob["obs_times"]).astype(np.int32)
    return pr, ob, ind


def host_threads():
    """every hardware thread this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    silently turn the CPU arm into a single-thread run"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def cpu_run(pr, ob, ind, B, threads=0):
    from oracle import c_port
    threads = threads or host_threads()
    t0 = time.perf_counter()
    out = c_port.dalton("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"][:B], 0.0, T_MAX, N_STEPS, pr["Q"], pr["R"],
                        pr["theta"][:B], ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"], n_threads=threads)
    return time.perf_counter() - t0, out


def cpu_baseline(target_seconds=10.0):
    """theta*steps/s of the C port with all host threads on a bounded sample (about `target_seconds` of CPU work)."""
    cores = host_threads()
    pr, ob, ind = cpu_problem(B_PER_GPU)
    dt, _ = cpu_run(pr, ob, ind, 2048)                      # calibration (also warms the thread pool)
    rate = 2048 * N_STEPS / dt
    Bs = int(min(B_PER_GPU, max(2048, rate * target_seconds / N_STEPS)))
    dt, _ = cpu_run(pr, ob, ind, Bs)
    return {"value": Bs * N_STEPS / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{Bs} of {B_PER_GPU} thetas x {N_STEPS} steps, FN dalton f64, C/OpenMP port of the reference "
                      f"algorithm (reference JAX unavailable), {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    pr, ob, ind = cpu_problem(B_PER_GPU)
    dt, _ = cpu_run(pr, ob, ind, 2048)
    rate = 2048 * N_STEPS / dt
    Bs = int(min(B_PER_GPU, max(2048, rate * 2.0 / N_STEPS)))          # ~2 s per step
    for _ in range(args.warmup):
        cpu_run(pr, ob, ind, Bs)
    times = [cpu_run(pr, ob, ind, Bs)[0] for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    value = Bs * N_STEPS / (ms * 1e-3)
    sample = f"{Bs} of {B_PER_GPU} thetas x {N_STEPS} steps per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "FitzHugh-Nagumo dalton log-likelihood, 65,536 thetas, n_steps=800, n_obs=41, f64 "
                               "(BASELINE configs[1]); bounded sample per step: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C/OpenMP port of the reference algorithm on all host threads; the reference's jit+vmap JAX-CPU "
                "path cannot run here (jax not installed, no network)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.stop, self.ok = [], set(), threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log("clock sampling unavailable:", e)
            self.max_mhz = None
        self.th = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        while not self.stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.ok:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.ok:
            self.th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rodeo_b200
    from rodeo_b200 import _host, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rodeo_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    dev = torch.device("cuda", local)
    B, N = args.thetas, N_STEPS

    # ---- synthetic inputs: each rank owns its own contiguous shard of the global theta batch (weak scaling)
    pr = workload(B, seed=rank)
    pr0 = P.fitz_problem(1, N, T_MAX, jitter=False)
    fn, kramer = rodeo_b200.models.fitzhugh_nagumo, rodeo_b200.interrogate.interrogate_kramer
    truth, _ = rodeo_b200.solve_mv(None, fn, pr0["W"], pr0["X0"][0], 0.0, T_MAX, N, kramer,
                                   prior_pars=(pr0["Q"], pr0["R"]), theta=pr0["theta"][0])
    ob = obs_for(pr, truth.cpu().numpy())

    pb = _host.Problem(None, fn, pr["W"], pr["X0"], 0.0, T_MAX, N, kramer, (pr["Q"], pr["R"]), None, None,
                       "standard", {"theta": pr["theta"]}, particle_offset=rank * B)
    pb.set_obs(ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    out = torch.empty((B,), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B,), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def step():
        rc = lib.rodeo_b200_dalton_f64(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                       _host.ptr(pb.x0), _host.ptr(pb.theta), None, _host.ptr(pb.obs_ind),
                                       _host.ptr(pb.obs_data), _host.ptr(pb.obs_weight), _host.ptr(pb.obs_var),
                                       _host.ptr(out), None, 0, pb.stream())
        _lib.check(rc, "dalton")
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()

    # ---- device-resident timed region: K steps, CUDA events on the launching stream, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = lib.rodeo_b200_launch_count()
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            step()
            e1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = lib.rodeo_b200_launch_count() - launches0
    step_ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    total_ms = torch.tensor([float(step_ms.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = world * B * N / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers (pinned), H2D + kernel + D2H inside the timed region
    h_x0 = torch.from_numpy(pr["X0"]).pin_memory()
    h_th = torch.from_numpy(pr["theta"]).pin_memory()
    h_out = torch.empty((B,), dtype=torch.float64).pin_memory()
    h_ind = np.ascontiguousarray(pb.obs_ind_host)
    h_y, h_D, h_Om = (np.ascontiguousarray(ob[k]) for k in ("obs_data", "obs_weight", "obs_var"))

    def e2e_step():
        rc = lib.rodeo_b200_dalton_f64_host(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                            ctypes.c_void_p(h_x0.data_ptr()), ctypes.c_void_p(h_th.data_ptr()),
                                            _host.ptr(h_ind), _host.ptr(h_y), _host.ptr(h_D), _host.ptr(h_Om),
                                            ctypes.c_void_p(h_out.data_ptr()))
        _lib.check(rc, "dalton_host")

    n_e2e = 0 if args.skip_e2e else args.steps
    for _ in range(0 if args.skip_e2e else 3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()                      # returns after the D2H copy has completed (stream synchronised inside)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * N * n_e2e / float(t_e2e.item()) if n_e2e else None
    h2d = int(h_x0.numel() * 8 + h_th.numel() * 8 + h_ind.nbytes + h_y.nbytes + h_D.nbytes + h_Om.nbytes)
    d2h = int(h_out.numel() * 8)
    same = bool(np.array_equal(h_out.numpy(), out.cpu().numpy(), equal_nan=True)) if n_e2e else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel: FP64-pipe bound, algorithmic flops / measured DFMA peak
    peak = ctypes.c_double(0.0)
    _lib.check(lib.rodeo_b200_fp64_peak_probe(5, ctypes.byref(peak)), "fp64 probe")
    kern_ms = float(np.mean(step_ms)) if world == 1 else ms_per_step
    achieved = FLOPS_PER_THETA_STEP * B * N / (kern_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic, fp64_instr = None, None
    tpath = os.path.join(ROOT, "profiles", "dalton_traffic.json")
    if os.path.exists(tpath):
        try:
            prof = json.load(open(tpath))
            traffic, fp64_instr = prof.get("dram_bytes_per_launch"), prof.get("fp64_instr_per_theta_step")
        except Exception:
            traffic = None
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": float(peak.value), "unit": "TFLOP/s",
        "frac": achieved / float(peak.value) if peak.value else None, "traffic": traffic,
        "peak_source": "DFMA micro-benchmark run in this process (rodeo_b200_fp64_peak_probe); "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flops_per_theta_step": FLOPS_PER_THETA_STEP,
        # hardware view: the kernel exploits the unit-triangular Q / unit-row W structure and executes fewer FP64
        # instructions than the dense count (hence frac > 1); this is the share of the FP64 pipe's issue slots it fills
        # (executed FP64 instructions per theta*step from the committed ncu source counters x measured rate / DFMA rate)
        "fp64_pipe_frac": (fp64_instr * B * N / (kern_ms * 1e-3)) / (float(peak.value) * 1e12 / 2.0)
                          if (fp64_instr and peak.value) else None,
        "executed_fp64_instr_per_theta_step": fp64_instr,
        "hbm_view": {"algorithmic_bytes_per_launch": int(B * (6 + 3 + 1) * 8),
                     "hbm_gbs_measured": peaks.get("hbm_gbs")},
    }

    cpu = cpu_baseline() if (world == 1 and not args.skip_cpu) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "FitzHugh-Nagumo dalton log-likelihood (BASELINE configs[1]): 65,536 thetas per GPU, "
                               "n_steps=800, t in [0,40], n_obs=41, interrogate_kramer, IBM sigma=0.1, float64",
                   "thetas_per_gpu": B, "n_steps": N, "n_obs": N_OBS, "parallelism": f"theta-sharded x{world}",
                   "l2": "256 MiB memset between timed steps (outside the event brackets)"},
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "rodeo_b200_dalton_f64_host (C ABI, pinned host buffers)", "matches_device_path": same},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--thetas", type=int, default=B_PER_GPU,
                    help="thetas per GPU (tuning experiments only; the benchmark configuration is the default)")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: skip the host-buffer e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
