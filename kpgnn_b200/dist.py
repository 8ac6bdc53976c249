"""Data-parallel plumbing for the K-hop path.

Graphs are independent units: every rank extracts, plans and aggregates its own contiguous shard of the global
batch, and the only exchange is ONE all-reduce (sum, then / world) of the flat fp32 gradient per step --
the one-process-per-GPU equivalent of the reference's PyG `DataParallel` (train_ZINC.py:90-91,181-189).
Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests); BatchNorm statistics stay per rank, as in the
reference's replicas.
"""
import torch
import torch.distributed as dist


def shard_bounds(num_items, rank, world):
    """Contiguous, equal-count shards (the last ranks get one item fewer when not divisible)."""
    base, rem = divmod(num_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_graphs(graphs, rank, world):
    lo, hi = shard_bounds(len(graphs), rank, world)
    return graphs[lo:hi]


class FlatGradients(object):
    """All parameter gradients as views into one contiguous fp32 buffer, so the step needs a single collective
    (and a single memset) instead of one per parameter."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero_(self):
        self.flat.zero_()

    def release(self):
        """Gather mode: let autograd hand over fresh gradient tensors (no per-parameter accumulate kernel) ..."""
        for p in self.params:
            p.grad = None

    def gather_(self):
        """... and pack them into the flat buffer with ONE concatenation after backward; parameters without a gradient
        contribute zeros.  Leaves every p.grad as a view of the flat buffer again (what the optimizer reads)."""
        pieces = []
        o = 0
        for p in self.params:
            g = p.grad
            pieces.append(g.reshape(-1) if g is not None else self.flat.new_zeros(p.numel()))
            o += p.numel()
        torch.cat(pieces, out=self.flat)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        return self.flat

    def allreduce_mean_(self, world=None):
        """Sum over ranks, divide by the world size (loss = per-rank mean => global mean for equal shards)."""
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.flat)
            self.flat.div_(world)
        return self.flat


class _RawCuda(object):
    """A float32 device buffer owned by the library, presented to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class PeerGradients(FlatGradients):
    """FlatGradients whose exchange is the library's own kernel over NVLink peer memory (csrc/peer.cu,
    include/kpgnn.h kp_peer_*): the local gradient is packed into a peer-visible block, every rank reads all blocks
    directly and writes the rank-ordered mean into `flat` (what the optimizer reads).  One kernel per step instead of
    {NCCL launch, divide, second graph}; no NCCL call inside the step, so the whole step is ONE CUDA graph.

    Set-up needs a process group for the one-time handle exchange only (any backend).  Raises when a peer block cannot
    be opened (no P2P path between the devices) -- the caller then keeps FlatGradients + the backend's all-reduce."""

    def __init__(self, params, group=None, barrier=True):
        from . import _lib
        import ctypes as C
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n_real = sum(p.numel() for p in self.params)
        n = (n_real + 3) & ~3                         # the kernel moves float4s; the pad floats stay zero
        self.n, self.n_real, self.device = n, n_real, dev
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        L = _lib.lib()
        with torch.cuda.device(dev):
            # Every rank takes part in the handle exchange even when its own allocation failed (payload None), and every
            # rank then sees the same list: either all raise here or none does -- no rank is left waiting in a collective.
            own, payload, err = C.c_void_p(), None, None
            try:
                if self.world > _lib.PEER_MAX:
                    raise RuntimeError("peer exchange supports up to %d ranks on one node" % _lib.PEER_MAX)
                nbytes = L.kp_peer_block_bytes(n)
                _lib.check(L.kp_peer_alloc(nbytes, C.byref(own)), "kp_peer_alloc")
                handle = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
                _lib.check(L.kp_peer_export(own, handle), "kp_peer_export")
                payload = (bytes(handle.raw), n, int(dev.index))
            except Exception as e:                      # noqa: BLE001 -- reported to every rank through the exchange
                err = e
            handles = [None] * self.world
            dist.all_gather_object(handles, payload, group=group)
            bad = [r for r, h in enumerate(handles) if h is None or h[1] != n]
            if bad:
                if own.value:
                    L.kp_peer_free(own)
                raise RuntimeError("peer gradient blocks unavailable on rank(s) %s%s" % (bad, ": %s" % err if err else ""))
            self.blocks = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.blocks.append(own.value)
                else:
                    ptr = C.c_void_p()
                    _lib.check(L.kp_peer_import(h[0], C.byref(ptr)), "kp_peer_import")
                    self.blocks.append(ptr.value)
            self._own = own.value
            vec_bytes = (n * 4 + 255) & ~255
            self._send_all = torch.as_tensor(_RawCuda(own.value + _lib.PEER_FLAG_BYTES, n), device=dev)
            self._recv_all = torch.as_tensor(_RawCuda(own.value + _lib.PEER_FLAG_BYTES + vec_bytes, n), device=dev)
            assert self._send_all.data_ptr() == own.value + _lib.PEER_FLAG_BYTES and self._send_all.dtype == torch.float32
            self.send, self.flat = self._send_all[:n_real], self._recv_all[:n_real]
            self.state = torch.zeros(_lib.PEER_CTAS + 1, dtype=torch.int32, device=dev)      # epochs | error
            d = _lib.PeerDesc()
            d.world, d.rank, d.n = self.world, self.rank, n
            for r in range(self.world):
                d.block[r] = self.blocks[r]
            d.out, d.epoch = None, self.state.data_ptr()
            d.error = self.state.data_ptr() + 4 * _lib.PEER_CTAS
            d.scale = 1.0 / self.world
            self.desc = d
        self._point_grads()
        # every rank must have opened every block before the first exchange.  A caller that follows the construction with
        # its own collective on all ranks (Trainer: the all-or-none agreement, which also covers a rank whose set-up
        # raised) passes barrier=False -- a barrier here would pair with that rank's agreement call and desynchronise.
        if barrier:
            dist.barrier(group=group)

    def _point_grads(self):
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def gather_(self):
        pieces = []
        for p in self.params:
            g = p.grad
            pieces.append(g.reshape(-1) if g is not None else self.flat.new_zeros(p.numel()))
        torch.cat(pieces, out=self.send)
        self._point_grads()
        return self.send

    def allreduce_mean_(self, world=None):
        from . import _lib
        import ctypes as C
        st = torch.cuda.current_stream(self.device)
        _lib.check(_lib.lib().kp_peer_allreduce_mean(C.byref(self.desc), C.c_void_p(st.cuda_stream)),
                   "kp_peer_allreduce_mean")
        return self.flat

    def close(self):
        """Unmaps the peers' blocks and frees this rank's (call on every rank, after a barrier: a peer may still be
        reading until its last exchange has returned).  The object is unusable afterwards."""
        from . import _lib
        L = _lib.lib()
        if getattr(self, "blocks", None) is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for r, ptr in enumerate(self.blocks):
                if r != self.rank and ptr:
                    L.kp_peer_release(ptr)
            self.send = self.flat = self._send_all = self._recv_all = None
            for p in self.params:
                p.grad = None
            L.kp_peer_free(self._own)
        self.blocks = None

    def check(self):
        """After a synchronisation point: did an exchange time out waiting for a peer?"""
        err = int(self.state[-1].item())
        if err:
            raise RuntimeError("peer gradient exchange timed out (%s)" %
                               ("waiting for the ranks' gradients" if err == 1 else "waiting for the readers"))
