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


class GatherPipeline:
    """Keeps the all-gather of the log-likelihoods off the critical path of a stream of batches.

    The collective of batch k runs on a side stream while the kernel of batch k+1 already runs on the compute stream
    (``depth`` local / gathered buffer pairs, reused round-robin).  Per batch:

        buf = pipe.local_buffer()      # (B_local,) tensor the kernel writes; waits, on the compute stream, for the
                                       # collective that last read this buffer
        ... launch the log-likelihood kernel into buf on the current stream ...
        slot = pipe.submit()           # enqueue the all-gather of buf behind the kernel, on the side stream
        ...
        full = pipe.result(slot)       # (world * B_local,) in rank order; the current stream waits for the collective

    Without an initialised process group (or world size 1) no collective is issued and ``result`` returns the local
    buffer.  Equal shard sizes only (pad with :func:`all_gather_loglik` otherwise).
    """

    def __init__(self, B_local, device, dtype=torch.float64, group=None, depth=2, collective=True):
        self.group = group
        self.world = (dist.get_world_size(group)
                      if (collective and dist.is_available() and dist.is_initialized()) else 1)
        self.depth, self.k = depth, 0
        self.local = [torch.zeros(B_local, dtype=dtype, device=device) for _ in range(depth)]
        self.cuda = torch.device(device).type == "cuda"
        if self.world > 1:
            self.gathered = [torch.empty(self.world * B_local, dtype=dtype, device=device) for _ in range(depth)]
            if self.cuda:
                self.comm = torch.cuda.Stream(device=device)
                self.ready = [torch.cuda.Event() for _ in range(depth)]
                self.done = [torch.cuda.Event() for _ in range(depth)]
                self.pending = [False] * depth

    def local_buffer(self):
        slot = self.k % self.depth
        if self.world > 1 and self.cuda and self.pending[slot]:
            torch.cuda.current_stream().wait_event(self.done[slot])      # its previous gather has read it
        return self.local[slot]

    def submit(self):
        slot = self.k % self.depth
        self.k += 1
        if self.world == 1:
            return slot
        if not self.cuda:
            dist.all_gather_into_tensor(self.gathered[slot], self.local[slot], group=self.group)
            return slot
        self.ready[slot].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ready[slot])
            dist.all_gather_into_tensor(self.gathered[slot], self.local[slot], group=self.group)
            self.done[slot].record(self.comm)
        self.pending[slot] = True
        return slot

    def result(self, slot):
        if self.world == 1:
            return self.local[slot]
        if self.cuda and self.pending[slot]:
            torch.cuda.current_stream().wait_event(self.done[slot])
        return self.gathered[slot]

    def drain(self):
        """the current stream waits for every collective issued so far"""
        if self.world > 1 and self.cuda:
            for slot in range(self.depth):
                if self.pending[slot]:
                    torch.cuda.current_stream().wait_event(self.done[slot])


def sharded_loglik(loglik_fn, theta, ode_init, group=None):
    """Run ``loglik_fn(theta_shard, ode_init_shard, particle_offset) -> (B_local,) tensor`` on this rank's shard
    and all-gather.  ``loglik_fn`` is typically a closure over ``rodeo_b200.inference.dalton`` / ``fenrir``."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(len(theta), rank, world)
    local = loglik_fn(theta[lo:hi], ode_init[lo:hi], lo)
    # a log-likelihood of host arrays comes back on the host (rodeo_b200.inference.dalton): NCCL gathers device tensors
    if world > 1 and dist.get_backend(group) == "nccl" and not local.is_cuda:
        local = local.to(torch.device("cuda", torch.cuda.current_device()))
    return all_gather_loglik(local, len(theta), group)
