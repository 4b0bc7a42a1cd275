"""Theta-batch sharding over the GPUs of one node: one process per GPU, ``torch.distributed`` for the plumbing.

Every theta (or (theta, seed) particle) is an independent filter -- there is no cross-theta term anywhere in
reference src/rodeo/solve.py / inference/*.py -- so the batch axis is split contiguously over ranks with NO
data-path collective.  The only exchange is the all-gather of the per-theta log-likelihoods (8 bytes each) after
the kernel; trajectories (solve_mv / solve_sim outputs) stay on the GPU that produced them.  Random streams are
keyed by the GLOBAL particle index (``particle_offset``), so results do not depend on the sharding.
"""
import torch
import torch.distributed as dist


def shard_bounds(B, rank, world):
    """Contiguous [lo, hi) of rank's shard; the first ``B % world`` ranks get one extra theta."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(x, rank, world):
    """Rows [lo, hi) of a batched array-like (leading axis = theta)."""
    lo, hi = shard_bounds(len(x), rank, world)
    return x[lo:hi]


def all_gather_loglik(local, B_total=None, group=None):
    """Concatenate every rank's per-theta log-likelihoods in rank order (NCCL on GPUs, gloo on CPU).

    Shards may differ by one element (uneven split): they are padded to a common length for the collective.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if all(s == m for s in sizes):
        out = torch.empty(world * m, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def sharded_loglik(loglik_fn, theta, ode_init, group=None):
    """Run ``loglik_fn(theta_shard, ode_init_shard, particle_offset) -> (B_local,) tensor`` on this rank's shard
    and all-gather.  ``loglik_fn`` is typically a closure over ``rodeo_b200.inference.dalton`` / ``fenrir``."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(len(theta), rank, world)
    local = loglik_fn(theta[lo:hi], ode_init[lo:hi], lo)
    return all_gather_loglik(local, len(theta), group)
