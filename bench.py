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


def cpu_baseline(target_seconds=10.0, return_outputs=False):
    """theta*steps/s of the C port with all host threads on a bounded sample (about `target_seconds` of CPU work)."""
    cores = host_threads()
    pr, ob, ind = cpu_problem(B_PER_GPU)
    dt, _ = cpu_run(pr, ob, ind, 2048)                      # calibration (also warms the thread pool)
    rate = 2048 * N_STEPS / dt
    Bs = int(min(B_PER_GPU, max(2048, rate * target_seconds / N_STEPS)))
    dt, out = cpu_run(pr, ob, ind, Bs)
    rec = {"value": Bs * N_STEPS / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{Bs} of {B_PER_GPU} thetas x {N_STEPS} steps, FN dalton f64, C/OpenMP port of the reference "
                     f"algorithm (scalar, dense, -ffp-contract=off; NOT the reference's JAX, which cannot run here), "
                     f"{dt:.2f} s"}
    if not return_outputs:
        return rec
    cargs = ("fitzhugh_nagumo", "kramer", pr["W"], pr["X0"][:Bs], 0.0, T_MAX, N_STEPS, pr["Q"], pr["R"],
             pr["theta"][:Bs], ob["obs_data"], ind, ob["obs_weight"], ob["obs_var"])
    return rec, out, cargs


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

    def __init__(self, index, period=0.004):
        self.samples, self.power, self.reasons, self.stop, self.ok = [], [], set(), threading.Event(), False
        self.period = period
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
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

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
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sm_mhz_min": float(np.min(self.samples)),
                "power_w_median": float(np.median(self.power)) if self.power else None}


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
N_COPIES = 64     # resident copies of a launch's inputs (X0, theta: 4.7 MB each); launches rotate over them


class Dalton:
    """one rank's dalton launcher: resident inputs, the C-ABI call, output into a caller-chosen buffer.

    L2 hygiene: the inputs of a launch are 4.7 MB, far below the 126 MB L2, so consecutive launches rotate over N_COPIES
    identical copies (300 MB in total): the copy a launch reads was last touched 64 launches and 300 MB of traffic
    earlier.  (The kernel is FP64-bound: a cold read of its inputs is < 1 us of a 0.74 ms launch.)"""

    def __init__(self, lib, pr, ob, lo, hi, offset, copies=N_COPIES):
        import rodeo_b200
        from rodeo_b200 import _host
        self.lib, self._host = lib, _host
        fn, kramer = rodeo_b200.models.fitzhugh_nagumo, rodeo_b200.interrogate.interrogate_kramer
        self.pb = _host.Problem(None, fn, pr["W"], pr["X0"][lo:hi], 0.0, T_MAX, N_STEPS, kramer, (pr["Q"], pr["R"]), None,
                                None, "standard", {"theta": pr["theta"][lo:hi]}, particle_offset=offset)
        self.pb.set_obs(ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
        self.B = hi - lo
        self.x0 = self.pb.x0.unsqueeze(0).repeat(copies, 1, 1, 1).contiguous()
        self.theta = self.pb.theta.unsqueeze(0).repeat(copies, 1, 1).contiguous()
        self.copies, self.k = copies, 0

    def __call__(self, out):
        import ctypes as C
        from rodeo_b200 import _lib
        pb, h = self.pb, self._host
        c = self.k % self.copies
        self.k += 1
        rc = self.lib.rodeo_b200_dalton_f64(C.byref(pb.c), h.ptr(pb.W), h.ptr(pb.Q), h.ptr(pb.R), h.ptr(self.x0[c]),
                                            h.ptr(self.theta[c]), None, h.ptr(pb.obs_ind), h.ptr(pb.obs_data),
                                            h.ptr(pb.obs_weight), h.ptr(pb.obs_var), h.ptr(out), None, 0, pb.stream())
        _lib.check(rc, "dalton")


def timed_steps(torch, dist, world, dev, launcher, pipe, steps):
    """K steps = K kernel launches, each followed (world > 1) by the all-gather of its log-likelihoods through the
    product's GatherPipeline (the collective of step k overlaps the kernel of step k+1; the last one is drained inside
    the timed region).  Returns (ms per step: wall of the whole region on the device, max over ranks; per-launch kernel
    ms from events around each launch)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for e0, e1 in kev:
        buf = pipe.local_buffer()
        e0.record()
        launcher(buf)
        e1.record()
        pipe.submit()
    pipe.drain()
    t1.record()
    barrier()
    total = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    return float(total.item()) / steps, np.array([e0.elapsed_time(e1) for e0, e1 in kev])


def run_ours(args):
    import torch
    import rodeo_b200
    from rodeo_b200 import _host, _lib, parallel

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
    steps, warm = args.steps, max(args.warmup, 3)

    # ---- synthetic inputs.  Weak scaling: rank r owns thetas [r B, (r+1) B) of a world*B batch (its own seed).
    pr = workload(B, seed=rank)
    pr0 = P.fitz_problem(1, N, T_MAX, jitter=False)
    fn, kramer = rodeo_b200.models.fitzhugh_nagumo, rodeo_b200.interrogate.interrogate_kramer
    truth, _ = rodeo_b200.solve_mv(None, fn, pr0["W"], pr0["X0"][0], 0.0, T_MAX, N, kramer,
                                   prior_pars=(pr0["Q"], pr0["R"]), theta=pr0["theta"][0])
    ob = obs_for(pr, truth.cpu().numpy())
    weak = Dalton(lib, pr, ob, 0, B, rank * B)
    pipe = parallel.GatherPipeline(B, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        weak(pipe.local_buffer())
        pipe.submit()
    pipe.drain()
    barrier()

    # ---- device-resident timed region
    launches0 = lib.rodeo_b200_launch_count()
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        ms_per_step, kern_ms = timed_steps(torch, dist, world, dev, weak, pipe, steps)
        t_wall = time.perf_counter() - t_wall0
    launches = lib.rodeo_b200_launch_count() - launches0
    value = world * B * N / (ms_per_step * 1e-3)
    out = pipe.local[(pipe.k - 1) % pipe.depth].clone()                     # this rank's result of the last step
    gather_ok = None
    if world > 1:
        # the gathered vector holds every rank's shard in rank order: this rank's slice is its own output, bitwise, and
        # every rank holds the same vector (compare a position-weighted checksum across ranks)
        full = pipe.result((pipe.k - 1) % pipe.depth)
        own = bool(torch.equal(full[rank * B:(rank + 1) * B], out))
        w = torch.arange(1, full.numel() + 1, dtype=torch.float64, device=dev)
        cs = torch.stack([(full * w).sum(), -(full * w).sum()])
        dist.all_reduce(cs, op=dist.ReduceOp.MAX)
        same = bool((cs[0] == -cs[1]).item())
        flag = torch.tensor([int(own and same and bool(torch.isfinite(full).all()))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(flag.item())
        if not gather_ok:
            raise SystemExit("bench.py: the gathered log-likelihoods do not match the per-rank outputs")

    # ---- the same weak-scaling step with the all-gather as the library's peer-store kernel on the side stream instead
    #      of NCCL's (GatherPipeline(peer=PeerGather)): reported next to the headline, which stays on NCCL
    weak_peer = None
    if world > 1:
        pgw = None
        try:
            pgw = parallel.PeerGather(world * B)
        except Exception as e:                                              # raised on every rank alike
            weak_peer = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        if pgw is not None:
            ok_w, t_w, same_w = 1, float("nan"), False
            try:
                ppipe = parallel.GatherPipeline(B, dev, peer=pgw)
                for _ in range(warm):
                    weak(ppipe.local_buffer()); ppipe.submit()
                ppipe.drain()
                t_w, _ = timed_steps(torch, dist, world, dev, weak, ppipe, steps)
                pgw.check()
                same_w = bool(torch.equal(ppipe.result((ppipe.k - 1) % ppipe.depth), full))
            except Exception as e:
                ok_w = 0
                weak_peer = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            agree = torch.tensor([float(ok_w), float(same_w)], dtype=torch.float64, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if bool(agree[0].item()):
                weak_peer = {"ms_per_step": t_w, "value": world * B * N / (t_w * 1e-3), "unit": UNIT,
                             "equals_nccl_gather_bitwise": bool(agree[1].item()),
                             "what": "GatherPipeline(peer=PeerGather): the gather of step k is one kernel of NVLink peer "
                                     "stores on the side stream while the dalton kernel of step k+1 runs"}
            elif weak_peer is None:
                weak_peer = {"unavailable": "failed on another rank"}
            try:
                barrier()
                pgw.close()
            except Exception:
                pass

    # ---- strong scaling at BASELINE configs[1]'s stated total: 65,536 thetas over `world` GPUs
    strong = None
    if args.thetas == B_PER_GPU:
        prg = workload(B_PER_GPU, seed=0)                                    # the global batch, identical on every rank
        obg = obs_for(prg, truth.cpu().numpy())
        lo, hi = parallel.shard_bounds(B_PER_GPU, rank, world)
        sh = Dalton(lib, prg, obg, lo, hi, lo)
        spipe = parallel.GatherPipeline(hi - lo, dev)
        for _ in range(warm):
            sh(spipe.local_buffer()); spipe.submit()
        spipe.drain()
        s_ms, s_kern = timed_steps(torch, dist, world, dev, sh, spipe, steps)
        rec = {"thetas_total": B_PER_GPU, "thetas_per_gpu": hi - lo, "ms_per_step": s_ms,
               "value": B_PER_GPU * N / (s_ms * 1e-3), "unit": UNIT, "kernel_ms": float(np.mean(s_kern)),
               "collective": "all-gather of the shard's log-likelihoods inside the step" if world > 1 else None}
        if world > 1:
            # one GPU doing the whole batch, measured in the same run (every rank times it; max over ranks)
            one = Dalton(lib, prg, obg, 0, B_PER_GPU, 0)
            opipe = parallel.GatherPipeline(B_PER_GPU, dev, collective=False)
            for _ in range(warm):
                one(opipe.local_buffer()); opipe.submit()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one(opipe.local_buffer()); opipe.submit()
            e1.record(); barrier()
            t1 = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
            dist.all_reduce(t1, op=dist.ReduceOp.MAX)
            full1 = opipe.local[(opipe.k - 1) % opipe.depth]
            fullN = spipe.result((spipe.k - 1) % spipe.depth)
            rec.update({"one_gpu_ms_per_step": float(t1.item()), "speedup_vs_one_gpu": float(t1.item()) / s_ms,
                        "sharded_equals_unsharded_bitwise": bool(torch.equal(full1, fullN))})
            rec["collective"] = ("nccl_pipeline: NCCL all-gather of the shard's log-likelihoods through GatherPipeline "
                                 "(side stream, overlapping the next launch; host-driven); graph_peer: a kernel of peer "
                                 "stores behind the dalton kernel, both in one CUDA graph per rank")
            # The same step as ONE CUDA graph per rank -- output memset, dalton kernel, and the all-gather as a kernel of
            # peer stores over NVLink (rodeo_b200.parallel.PeerGather) -- replayed back to back: no per-step host work.
            gp, t_gp, ok_local, same_gp = {}, float("nan"), 1, False
            pg = None
            try:
                pg = parallel.PeerGather(B_PER_GPU)
            except Exception as e:                                      # raised on every rank alike (collective agreement)
                gp = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            if pg is not None:
                try:
                    gbuf = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
                    n_calls = 0
                    for _ in range(3):
                        sh(gbuf); pg.gather(gbuf, device_epoch=True); n_calls += 1
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):   # (other threads: NCCL watchdog, clock sampler)
                        sh(gbuf); pg.gather(gbuf, device_epoch=True)
                    for _ in range(3):
                        graph.replay(); n_calls += 1
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        graph.replay()
                    e1.record(); torch.cuda.synchronize()
                    n_calls += steps
                    t_gp = e0.elapsed_time(e1) / steps
                    pg.check()
                    same_gp = bool(torch.equal(pg.result(n_calls), fullN))
                except Exception as e:
                    ok_local = 0
                    gp = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
                agree = torch.tensor([float(ok_local), float(same_gp)], dtype=torch.float64, device=dev)
                dist.all_reduce(agree, op=dist.ReduceOp.MIN)
                tg = torch.tensor([t_gp if ok_local else 0.0], dtype=torch.float64, device=dev)
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
                if bool(agree[0].item()):
                    gp = {"ms_per_step": float(tg.item()), "value": B_PER_GPU * N / (float(tg.item()) * 1e-3), "unit": UNIT,
                          "speedup_vs_one_gpu": float(t1.item()) / float(tg.item()),
                          "equals_nccl_gather_bitwise": bool(agree[1].item()),
                          "what": "one CUDA graph per rank (output memset + dalton kernel + peer-store all-gather kernel "
                                  "over NVLink, CUDA-IPC mapped regions), replayed back to back; device time, max over ranks"}
                elif not gp:
                    gp = {"unavailable": "failed on another rank"}
                try:
                    barrier()
                    pg.close()
                except Exception:
                    pass
            rec["graph_peer"] = gp
            # the record's headline figures are those of the faster path; the other one stays next to it
            rec["nccl_pipeline"] = {"ms_per_step": s_ms, "value": B_PER_GPU * N / (s_ms * 1e-3),
                                    "speedup_vs_one_gpu": float(t1.item()) / s_ms}
            if "ms_per_step" in gp and gp.get("equals_nccl_gather_bitwise") and gp["ms_per_step"] < s_ms:
                rec.update({"ms_per_step": gp["ms_per_step"], "value": gp["value"],
                            "speedup_vs_one_gpu": gp["speedup_vs_one_gpu"], "path": "graph_peer"})
            else:
                rec["path"] = "nccl_pipeline"
        strong = rec

    # ---- BASELINE configs[4] at its stated total: 262,144 particles of one pseudo-marginal iteration (solve_sim +
    #      interrogate_chkrebtii + Gaussian observation log-likelihood, ONE fused kernel, no trajectories written) over
    #      the `world` GPUs, the per-particle log-likelihoods all-gathered
    strong_c5 = None
    if args.thetas == B_PER_GPU and not args.skip_configs:
        import functools
        P5 = 262144
        lo, hi = parallel.shard_bounds(P5, rank, world)
        pr5 = workload(P5, seed=5)
        ob5 = obs_for(pr5, truth.cpu().numpy())
        chk = functools.partial(rodeo_b200.interrogate.interrogate_chkrebtii, kalman_type="standard")
        X5 = torch.as_tensor(pr5["X0"][lo:hi], device=dev)
        th5 = torch.as_tensor(pr5["theta"][lo:hi], device=dev)
        Y5 = torch.as_tensor(ob5["obs_data"][:, :, 0], device=dev)
        g5 = torch.empty(P5, dtype=torch.float64, device=dev) if world > 1 else None

        def c5_step(k):
            ll = rodeo_b200.solve_sim_loglik(np.array([9, k], dtype=np.uint32), fn, pr5["W"], X5, 0.0, T_MAX, N,
                                             chk, prior_pars=(pr5["Q"], pr5["R"]), theta=th5, obs_data=Y5,
                                             obs_times=ob5["obs_times"], noise_sd=float(np.sqrt(0.005)),
                                             _particle_offset=lo)
            return parallel.all_gather_loglik(ll, P5, out=g5)
        for k in range(2):
            full5 = c5_step(k)
        barrier()
        n5 = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n5):
            full5 = c5_step(10 + k)
        e1.record()
        barrier()
        t5 = torch.tensor([e0.elapsed_time(e1) / n5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        strong_c5 = {"particles_total": P5, "particles_per_gpu": hi - lo, "n_steps": N, "ms_per_iteration": float(t5.item()),
                     "value": P5 * N / (float(t5.item()) * 1e-3), "unit": UNIT,
                     "finite": bool(torch.isfinite(full5).all().item()),
                     "what": "rodeo_b200.solve_sim_loglik (fused solve_sim + chkrebtii + Gaussian obs log-lik, no Xt) per "
                             "shard + all-gather of the log-likelihoods"}
        del X5, th5, full5
        torch.cuda.empty_cache()

    # ---- end to end through the C ABI with host buffers (pinned), H2D + kernel + D2H inside the timed region
    h_x0 = torch.from_numpy(pr["X0"]).pin_memory()
    h_th = torch.from_numpy(pr["theta"]).pin_memory()
    h_out = torch.empty((B,), dtype=torch.float64).pin_memory()
    pb = weak.pb
    h_ind = np.ascontiguousarray(pb.obs_ind_host)
    h_y, h_D, h_Om = (np.ascontiguousarray(ob[k]) for k in ("obs_data", "obs_weight", "obs_var"))

    def e2e_step():
        rc = lib.rodeo_b200_dalton_f64_host(ctypes.byref(pb.c), _host.ptr(pb.W), _host.ptr(pb.Q), _host.ptr(pb.R),
                                            ctypes.c_void_p(h_x0.data_ptr()), ctypes.c_void_p(h_th.data_ptr()),
                                            _host.ptr(h_ind), _host.ptr(h_y), _host.ptr(h_D), _host.ptr(h_Om),
                                            ctypes.c_void_p(h_out.data_ptr()))
        _lib.check(rc, "dalton_host")

    n_e2e = 0 if args.skip_e2e else steps
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

    # ---- the call a rodeo user makes: the Python drop-in with NumPy inputs (pageable H2D, per-call allocation), result
    #      copied back to NumPy
    e2e_py = None
    if not args.skip_e2e:
        kw = dict(prior_pars=(pr["Q"], pr["R"]), theta=pr["theta"], **ob)

        def py_step():
            return rodeo_b200.inference.dalton(None, fn, pr["W"], pr["X0"], 0.0, T_MAX, N, kramer, **kw).cpu().numpy()
        for _ in range(2):
            r_py = py_step()
        barrier()
        n_py = max(3, min(steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_py):
            r_py = py_step()
        t_py = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_py, op=dist.ReduceOp.MAX)
        e2e_py = {"value": world * B * N * n_py / float(t_py.item()), "unit": UNIT,
                  "api": "rodeo_b200.inference.dalton(NumPy inputs) -> .cpu().numpy()",
                  "matches_device_path": bool(np.array_equal(r_py, out.cpu().numpy(), equal_nan=True))}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- sustained: >= 2 s of back-to-back launches (no flush, no host work in between), clocks sampled throughout
    sustained = None
    if world == 1 and not args.skip_sustained:
        buf = pipe.local_buffer()
        n_sus = int(2.2 / (float(np.mean(kern_ms)) * 1e-3)) + 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with ClockSampler(local, period=0.05) as sclk:
            e0.record()
            for _ in range(n_sus):
                weak(buf)
            e1.record()
            torch.cuda.synchronize()
        sus_ms = e0.elapsed_time(e1) / n_sus
        sustained = {"launches": n_sus, "seconds": e0.elapsed_time(e1) * 1e-3, "ms_per_step": sus_ms,
                     "value": B * N / (sus_ms * 1e-3), "unit": UNIT, "clocks": sclk.summary(),
                     "note": "back-to-back launches for >= 2 s with no host work in between (inputs rotate over the 64 "
                             "copies as in the burst figure)"}

    # ---- roofline of the dominant (only) kernel: FP64-pipe bound
    peak = ctypes.c_double(0.0)
    _lib.check(lib.rodeo_b200_fp64_peak_probe(5, ctypes.byref(peak)), "fp64 probe")
    k_ms = float(np.mean(kern_ms))
    achieved = FLOPS_PER_THETA_STEP * B * N / (k_ms * 1e-3) / 1e12
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
    pipe_frac = ((fp64_instr * B * N / (k_ms * 1e-3)) / (float(peak.value) * 1e12 / 2.0)
                 if (fp64_instr and peak.value) else None)
    # Two views of the same launch time.  SURVEY 8(d)'s yardstick is the DENSE algorithmic count (963 flop per theta*step,
    # no structure exploited): on it the kernel reads 2.0 x the measured DFMA peak, because it skips the multiplications by
    # the exact 0/1 entries of the unit-triangular Q and the unit-row W and conditions on an observation by two scalar
    # updates -- a fraction above 1 says the yardstick is wrong for this kernel, not that work is skipped.
    # `achieved` / `frac` are therefore the hardware view: FP64 instructions the kernel actually executes (ncu source
    # counters of the same command, profiles/dalton_traffic.json) x 2 flop / measured launch time, against the DFMA peak
    # measured in this process; the dense figures stay next to them as achieved_dense / frac_dense.
    exec_tflops = (2.0 * fp64_instr * B * N / (k_ms * 1e-3) / 1e12) if fp64_instr else None
    roofline = {
        "bound": "fp64", "achieved": exec_tflops if exec_tflops is not None else achieved, "peak": float(peak.value),
        "unit": "TFLOP/s", "frac": pipe_frac if pipe_frac is not None else (achieved / float(peak.value) if peak.value else None),
        "traffic": traffic,
        "achieved_is": "executed FP64 instructions x 2 flop / s" if exec_tflops is not None else "dense algorithmic flops / s",
        "achieved_dense": achieved,
        "frac_dense": achieved / float(peak.value) if peak.value else None,
        "frac_fp64_pipe": pipe_frac,
        "kernel_ms": k_ms,
        "peak_source": "DFMA micro-benchmark run in this process (rodeo_b200_fp64_peak_probe, private stream); "
                       "MEASURED_PEAKS.json has no FP64 figure; spec estimate 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
        "algorithmic_flops_per_theta_step": FLOPS_PER_THETA_STEP,
        "executed_fp64_instr_per_theta_step": fp64_instr,
        "hbm_view": {"algorithmic_bytes_per_launch": int(B * (6 + 3 + 1) * 8),
                     "hbm_gbs_measured": peaks.get("hbm_gbs")},
    }

    # ---- CPU baseline and parity of the TIMED launch: every theta against the C port (float64) of the cpu_baseline
    #      leg, judged against the same port in long double (tests/noise_floor.py)
    cpu, parity = None, None
    if world == 1 and not args.skip_cpu:
        cpu, cpu_out, cpu_args = cpu_baseline(return_outputs=True)
        nb_ = min(len(cpu_out), B)
        if nb_ > 0:
            import noise_floor as NF
            from oracle import c_port
            exact = c_port.dalton_ld(*cpu_args, n_threads=host_threads())
            parity = NF.gate(out.cpu().numpy()[:nb_], cpu_out[:nb_], exact[:nb_])
            parity["against"] = ("%d of the %d thetas of the timed launch; oracle = C port of the reference algorithm "
                                 "(float64), exact = the same port in x87 long double" % (nb_, B))

    configs = None
    if world == 1 and not args.skip_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        torch.cuda.empty_cache()
        configs = bench_configs.run("C1,C3,C4,C5", reps=3, quiet=True)
        cpath = os.path.join(ROOT, "profiles", "config_traffic.json")
        if os.path.exists(cpath):
            extra = json.load(open(cpath))
            for c in configs:
                c.update(extra.get(c["config"].split()[0], {}))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "FitzHugh-Nagumo dalton log-likelihood (BASELINE configs[1]): 65,536 thetas per GPU, "
                               "n_steps=800, t in [0,40], n_obs=41, interrogate_kramer, IBM sigma=0.1, float64",
                   "thetas_per_gpu": B, "n_steps": N, "n_obs": N_OBS, "parallelism": f"theta-sharded x{world}",
                   "l2": "inputs larger than L2: launches rotate over 64 resident copies of (X0, theta), 300 MB in total",
                   "step": "one dalton launch over the rank's batch" + (
                       " + NCCL all-gather of its log-likelihoods through rodeo_b200.parallel.GatherPipeline "
                       "(the collective of step k overlaps the kernel of step k+1; drained inside the timed region)"
                       if world > 1 else "")},
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "rodeo_b200_dalton_f64_host (C ABI, pinned host buffers)", "matches_device_path": same},
        "e2e_python": e2e_py, "weak_peer_gather": weak_peer, "strong": strong, "strong_c5": strong_c5, "sustained": sustained, "configs": configs,
        "gather_matches_rank_outputs": gather_ok,
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--skip-configs", action="store_true", help="skip the C1/C3/C4/C5 records")
    ap.add_argument("--skip-sustained", action="store_true", help="skip the >= 2 s back-to-back run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
