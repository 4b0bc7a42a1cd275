"""Worker of tests/test_multi_gpu.py: one rank per GPU (torchrun), NCCL.  The real kernels on every rank's shard, the
product's gather helpers, and the checks: sharded == unsharded bitwise, == oracle, every rank holds the same vector."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import functools
    import problems as P
    import rodeo_b200 as rb
    from rodeo_b200 import parallel
    from oracle import rodeo_oracle as orc
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, N, tm = 1000 + 3, 120, 6.0                                # uneven split over 2, 4 and 8 ranks
    pr = P.fitz_problem(B, n_steps=N, t_max=tm, seed=51)
    ob = P.fitz_obs(pr, None, n_obs=7)
    kr = rb.interrogate.interrogate_kramer
    fn = rb.models.fitzhugh_nagumo

    def loglik(theta, x0, offset):
        return rb.inference.dalton(None, fn, pr["W"], x0, 0.0, tm, N, kr, prior_pars=(pr["Q"], pr["R"]), theta=theta,
                                   _particle_offset=offset, **ob)

    full = parallel.sharded_loglik(loglik, pr["theta"], pr["X0"])         # all-gather with padding (uneven shards)
    alone = loglik(pr["theta"], pr["X0"], 0)                              # this GPU, the whole batch
    assert full.is_cuda and not alone.is_cuda        # host arrays in: the unsharded result comes back on the host
    assert full.shape == (B,) and torch.equal(full.cpu(), alone), "sharded != unsharded"
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, tm, N, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    err = np.max(np.abs(full.cpu().numpy() - want) / np.maximum(1, np.abs(want)))
    assert err < 1e-10, err

    # the pipelined gather the benchmark uses: even shards, three batches in flight over two buffer pairs
    Be = 64 * world
    pr2 = P.fitz_problem(Be, n_steps=N, t_max=tm, seed=52)
    lo, hi = parallel.shard_bounds(Be, rank, world)
    pipe = parallel.GatherPipeline(hi - lo, dev)
    ref = loglik(pr2["theta"], pr2["X0"], 0).to(dev)
    slots = []
    for k in range(3):
        buf = pipe.local_buffer()
        buf.copy_(loglik(pr2["theta"][lo:hi], pr2["X0"][lo:hi], lo) + float(k))
        slots.append(pipe.submit())
        if k >= 1:                                                        # consume batch k-1 while batch k is in flight
            assert torch.equal(pipe.result(slots[k - 1]), ref + float(k - 1))
    assert torch.equal(pipe.result(slots[2]), ref + 2.0)

    # the peer-memory all-gather (one kernel of peer stores over NVLink instead of the NCCL collective): uneven shards,
    # many calls so that both slots and the flag epochs are reused, every result equal to NCCL's bitwise
    pg = parallel.PeerGather(B, None)
    lo_u, hi_u = parallel.shard_bounds(B, rank, world)
    mine = full[lo_u:hi_u].clone()
    for k in range(40):
        loc = mine + float(k)
        got = pg.gather(loc)
        want_k = parallel.all_gather_loglik(loc, B)
        assert torch.equal(got, want_k), f"peer gather != NCCL gather at call {k}"
    assert torch.equal(parallel.sharded_loglik(loglik, pr["theta"], pr["X0"], peer=pg), full)     # the helper's peer path
    pg.check()
    dist.barrier()
    pg.close()
    # fewer elements than ranks: some shards are empty
    tiny = parallel.PeerGather(max(1, world - 1), None)
    tl, th_ = parallel.shard_bounds(max(1, world - 1), rank, world)
    vals = torch.arange(tl, th_, dtype=torch.float64, device=dev) + 0.5
    got = tiny.gather(vals)
    assert torch.equal(got, torch.arange(max(1, world - 1), dtype=torch.float64, device=dev) + 0.5)
    tiny.check()
    dist.barrier()
    tiny.close()

    # solve_sim: random streams are keyed by the GLOBAL particle index, so a shard reproduces its rows of the whole
    chk = functools.partial(rb.interrogate.interrogate_chkrebtii, kalman_type="standard")
    key = np.array([5, 7], dtype=np.uint32)
    xs = rb.solve_sim(key, fn, pr2["W"], pr2["X0"][lo:hi], 0.0, tm, N, chk, prior_pars=(pr2["Q"], pr2["R"]),
                      theta=pr2["theta"][lo:hi], _particle_offset=lo)
    xa = rb.solve_sim(key, fn, pr2["W"], pr2["X0"], 0.0, tm, N, chk, prior_pars=(pr2["Q"], pr2["R"]),
                      theta=pr2["theta"])
    assert torch.equal(xs, xa[lo:hi]), "solve_sim shard != rows of the whole"
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} dalton err vs oracle {err:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
