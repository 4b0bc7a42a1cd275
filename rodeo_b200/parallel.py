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

    def __init__(self, B_local, device, dtype=torch.float64, group=None, depth=2, collective=True, peer=None):
        """``peer``: a :class:`PeerGather` for ``world * B_local`` elements -- the collective on the side stream is then its
        kernel of peer stores instead of NCCL's all-gather (depth must stay 2: the peer regions have two slots)."""
        self.group = group
        self.peer = peer
        if peer is not None and depth != 2:
            raise ValueError("GatherPipeline with a PeerGather needs depth == 2")
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
            if self.peer is not None:
                self.gathered[slot] = self.peer.gather(self.local[slot])     # a view of the peer region's slot
            else:
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


def sharded_loglik(loglik_fn, theta, ode_init, group=None, peer=None):
    """Run ``loglik_fn(theta_shard, ode_init_shard, particle_offset) -> (B_local,) tensor`` on this rank's shard
    and all-gather.  ``loglik_fn`` is typically a closure over ``rodeo_b200.inference.dalton`` / ``fenrir``.
    ``peer``: a :class:`PeerGather` built for ``len(theta)`` -- the gather is then its one kernel of peer stores instead of
    the NCCL collective, and the result a view of its buffer (valid until the next-but-one gather)."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_bounds(len(theta), rank, world)
    local = loglik_fn(theta[lo:hi], ode_init[lo:hi], lo)
    # a log-likelihood of host arrays comes back on the host (rodeo_b200.inference.dalton): NCCL gathers device tensors
    if world > 1 and dist.get_backend(group) == "nccl" and not local.is_cuda:
        local = local.to(torch.device("cuda", torch.cuda.current_device()))
    if peer is not None and world > 1:
        return peer.gather(local.contiguous())
    return all_gather_loglik(local, len(theta), group)


class _DevPtr:
    """a raw device allocation as a __cuda_array_interface__ object (so that torch can view it without copying)"""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerGather:
    """All-gather of the per-theta log-likelihoods through peer memory instead of NCCL (one process per GPU, one node).

    Every rank owns a region (two slots of ``B_total`` float64 + flags) allocated by the library and mapped into the
    other ranks with CUDA IPC; ``gather(local)`` is ONE kernel on the current stream that stores the rank's shard into
    every region over NVLink / NVSwitch, publishes a flag per peer and waits -- bounded -- for the peers' flags
    (rodeo_b200/csrc/abi_peer.cu).  It returns a view of the rank's own slot: consume it on the same stream before the
    next ``gather``.  Construction is collective (handles are exchanged through ``torch.distributed``) and raises if the
    mappings cannot be made; callers fall back to :func:`all_gather_loglik`.  ``check()`` synchronises and raises if a
    peer's flag ever failed to arrive.
    """

    def __init__(self, B_total, group=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self._lib = _lib
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.B_total = int(B_total)
        self.lo, self.hi = shard_bounds(self.B_total, self.rank, self.world)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = self.lib.rodeo_b200_peer_region_bytes(self.B_total, self.world)
        if nbytes == 0:
            raise ValueError("PeerGather: world size out of range")
        own, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.rodeo_b200_peer_alloc(nbytes, ctypes.byref(own), ctypes.cast(handle, ctypes.c_void_p)),
                   "peer_alloc")
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs, self.opened, err = [None] * self.world, [], None
        for r in range(self.world):
            if r == self.rank:
                self.ptrs[r] = self.own
                continue
            p = ctypes.c_void_p()
            hb = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            rc = self.lib.rodeo_b200_peer_open(ctypes.cast(hb, ctypes.c_void_p), ctypes.byref(p))
            if rc != 0:
                err = self.lib.rodeo_b200_last_error().decode(errors="replace")
                break
            self.ptrs[r] = p.value
            self.opened.append(p.value)
        # every rank must agree before anyone stores into a peer
        ok = torch.tensor([0 if err else 1], device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise _lib.RodeoError(f"PeerGather: CUDA IPC mapping failed on some rank ({err})")
        self.regions = (ctypes.c_void_p * self.world)(*self.ptrs)
        self.status = torch.zeros(2, dtype=torch.int32, device=self.dev)       # [0] failure flag, [1] device-side epoch
        self.data = torch.as_tensor(_DevPtr(self.own, 2 * self.B_total), device=self.dev).view(2, self.B_total)
        self.epoch = 0
        dist.barrier(group=group)

    def gather(self, local, device_epoch=False):
        """local: this rank's (hi - lo,) float64 CUDA tensor -> (B_total,) view of the gathered vector.

        ``device_epoch=True`` keeps the call counter on the device so that the launch can be captured in a CUDA graph
        and replayed (every rank must then use it for every call); after ``n`` calls / replays in total the result is
        ``self.data[n & 1]`` (:meth:`result`)."""
        import ctypes
        if local.numel() != self.hi - self.lo or local.dtype != torch.float64 or not local.is_cuda:
            raise ValueError("PeerGather.gather: the rank's shard as a float64 CUDA tensor")
        if not local.is_contiguous():
            raise ValueError("PeerGather.gather: contiguous shard expected")
        if not device_epoch:
            self.epoch += 1
        rc = self.lib.rodeo_b200_peer_allgather_f64(
            ctypes.c_void_p(local.data_ptr()), local.numel(), self.lo, self.B_total, self.rank, self.world,
            ctypes.cast(self.regions, ctypes.c_void_p), 0 if device_epoch else self.epoch,
            ctypes.c_void_p(self.status.data_ptr() + 4), 0, ctypes.c_void_p(self.status.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check(rc, "peer_allgather")
        return None if device_epoch else self.data[self.epoch & 1]

    def result(self, n_calls):
        """the gathered vector after ``n_calls`` calls in total (device-epoch mode)"""
        return self.data[n_calls & 1]

    def check(self):
        if int(self.status[0].item()) != 0:
            raise self._lib.RodeoError("PeerGather: a peer's flag did not arrive (a rank missed a gather call?)")

    def close(self):
        for p in getattr(self, "opened", []):
            self.lib.rodeo_b200_peer_close(p)
        self.opened = []
        if getattr(self, "own", None):
            torch.cuda.synchronize()
            self.data = None                              # the view below dies with the region
            self.lib.rodeo_b200_peer_free(self.own)
            self.own = None
