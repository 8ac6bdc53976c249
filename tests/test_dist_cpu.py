"""CPU, world_size 2, gloo: the data-parallel host logic (sharding + single flat-gradient all-reduce) gives the
same gradients as one process over the whole batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kpgnn_b200.dist import FlatGradients, shard_bounds, shard_graphs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    X, y = torch.randn(16, 6, generator=g), torch.randn(16, generator=g)
    lo, hi = shard_bounds(16, rank, world)
    m = _model()
    fg = FlatGradients(m.parameters())
    fg.zero_()
    loss = (m(X[lo:hi]).squeeze() - y[lo:hi]).abs().mean()
    loss.backward()
    fg.allreduce_mean_()
    if rank == 0:
        torch.save(fg.flat.clone(), out)
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(1)
    X, y = torch.randn(16, 6, generator=g), torch.randn(16, generator=g)
    m = _model()
    fg = FlatGradients(m.parameters())
    (m(X).squeeze() - y).abs().mean().backward()
    assert torch.allclose(got, fg.flat, atol=1e-6)


def test_gather_mode_equals_view_mode():
    g = torch.Generator().manual_seed(2)
    X, y = torch.randn(8, 6, generator=g), torch.randn(8, generator=g)
    a, b = _model(), _model()
    fa, fb = FlatGradients(a.parameters()), FlatGradients(b.parameters())
    fa.zero_()
    (a(X).squeeze() - y).abs().mean().backward()
    fb.release()
    (b(X).squeeze() - y).abs().mean().backward()
    fb.gather_()
    assert torch.equal(fa.flat, fb.flat)
    assert all(p.grad.data_ptr() >= fb.flat.data_ptr() for p in b.parameters())


def test_shards_are_contiguous_and_cover():
    items = list(range(11))
    parts = [shard_graphs(items, r, 4) for r in range(4)]
    assert sum(parts, []) == items
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
