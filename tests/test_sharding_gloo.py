"""CPU, world_size 2, gloo: the N > 1 host path (contiguous theta shards, all-gather of per-theta log-likelihoods).

The per-shard compute here is the oracle (the product has no CPU path); what is under test is the sharding /
gather logic of rodeo_b200.parallel, i.e. that sharded == unsharded regardless of the split.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import problems as P
    from oracle import rodeo_oracle as orc
    from rodeo_b200 import parallel
    pr = P.fitz_problem(B, n_steps=40, t_max=2.0, seed=3)
    ob = P.fitz_obs(pr, None, n_obs=3)

    def local(theta, x0, offset):
        ll = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], x0, 0.0, 2.0, 40, orc.interrogate_kramer,
                        (pr["Q"], pr["R"]), theta, ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
        return torch.from_numpy(ll) + 0.0 * offset

    full = parallel.sharded_loglik(local, pr["theta"], pr["X0"])
    # the pipelined gather (CPU tensors: the collectives run in program order): three batches over two buffer pairs
    pipe = parallel.GatherPipeline(4, "cpu")
    for k in range(3):
        pipe.local_buffer().copy_(torch.arange(4, dtype=torch.float64) + 10 * rank + 100 * k)
        slot = pipe.submit()
        want = torch.cat([torch.arange(4, dtype=torch.float64) + 10 * r + 100 * k for r in range(world)])
        assert torch.equal(pipe.result(slot), want)
    pipe.drain()
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])          # even and uneven split
def test_sharded_equals_unsharded(B):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import problems as P
    from oracle import rodeo_oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pr = P.fitz_problem(B, n_steps=40, t_max=2.0, seed=3)
    ob = P.fitz_obs(pr, None, n_obs=3)
    want = orc.dalton(orc.MODELS["fitzhugh_nagumo"], pr["W"], pr["X0"], 0.0, 2.0, 40, orc.interrogate_kramer,
                      (pr["Q"], pr["R"]), pr["theta"], ob["obs_data"], ob["obs_times"], ob["obs_weight"], ob["obs_var"])
    assert got.shape == (B,) and np.array_equal(got, want)


def test_shard_bounds_cover_the_batch():
    from rodeo_b200.parallel import shard_bounds
    for B in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
