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
