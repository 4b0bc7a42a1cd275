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


def all_gather_loglik(local, B_total=None, group=None, out=None):
    """Concatenate every rank's per-theta log-likelihoods in rank order (NCCL on GPUs, gloo on CPU).

    With ``B_total`` (the global batch size) the shard sizes follow from :func:`shard_bounds`: no size exchange and
    no host synchronisation -- one collective, enqueued on the current stream.  An even split is a single
    ``all_gather_into_tensor`` (into ``out`` when given, so a steady-state caller allocates nothing); an uneven one
    (shards differ by one element) pads to the common length.  Without ``B_total`` every rank must hold the same
    number of elements.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.contiguous()
    if B_total is None:
        B_total = world * local.numel()
    sizes = [hi - lo for lo, hi in (shard_bounds(B_total, r, world) for r in range(world))]
    if local.numel() != sizes[dist.get_rank(group)]:
        raise ValueError(f"rank holds {local.numel()} log-likelihoods, shard_bounds({B_total}) gives "
                         f"{sizes[dist.get_rank(group)]}")
    m = max(sizes)
    if m == min(sizes):
        if out is None:
            out = torch.empty(world * m, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    buf = torch.empty(world * m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m:r * m + s] for r, s in enumerate(sizes)])


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
