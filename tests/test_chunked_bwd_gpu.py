"""Node-range (chunked) backward (include/kpgnn.h kp_agg_desc.node_base, kpgnn_b200/ops.py agg_backward): the backward
of a large batch run over ranges of whole graphs through ONE slice-sized hand-over workspace must give the gradients of
the single call -- dX and dP bit for bit (same kernels, same per-node order), table / theta gradients up to the
re-association of the chunk sums -- and calls that the packed-math kernels do not serve must fall back to the single call."""
import pytest
import torch

from kpgnn_b200 import synth
from tests.util import RTOL, collate, rel_err

pytestmark = pytest.mark.gpu


def _grads(b, K, d, act, fuse, tables, eps_on, chunk_bytes):
    from kpgnn_b200 import ops
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    N = b["num_nodes"]
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, K, d, generator=g).to(dev).requires_grad_(True)
    P = torch.randn(N, K, d, generator=g).to(dev).requires_grad_(True)
    T0 = torch.randn(5, d, generator=g).to(dev).requires_grad_(tables)
    Tk = torch.randn(52, d, generator=g).to(dev).requires_grad_(tables)
    th = torch.softmax(torch.randn(K, d, generator=g), 0).to(dev).requires_grad_(fuse)
    eps = torch.tensor([0.25], device=dev, requires_grad=True) if eps_on else None
    gy = torch.randn((N, d) if fuse else (N, K, d), generator=g).to(dev)
    plan, k = get_plan(ei, ea, N)
    old = (ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES)
    ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES = (None, old[1]) if chunk_bytes is None else (0, chunk_bytes)
    try:
        chunks = ops.backward_chunks(plan, K, d)
        out = ops.khop_aggregate(x, plan, K, P=P, T0=T0 if tables else None, Tk=Tk if tables else None,
                                 theta=th if fuse else None, eps=eps, act=act, fuse=fuse)
        out.backward(gy)
    finally:
        ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES = old
    leaves = {"dX": x.grad, "dP": P.grad}
    if tables:
        leaves.update(dT0=T0.grad, dTk=Tk.grad)
    if fuse:
        leaves["dtheta"] = th.grad
    if eps_on:
        leaves["deps"] = eps.grad
    return out.detach(), leaves, chunks


@pytest.mark.parametrize("act,fuse,tables", [("gelu", True, True), ("relu", False, True), ("gelu", False, False),
                                             ("gelu", True, False)])
def test_chunked_backward_equals_single_call(lib, act, fuse, tables):
    from kpgnn_b200.ops import ACT_GELU, ACT_RELU
    K, d = 8, 104
    b = collate(synth.zinc_like_graphs(96, seed=11), (K, 50, 6, 3, 50, 50, "spd"))
    a = {"gelu": ACT_GELU, "relu": ACT_RELU}[act]
    o1, g1, c1 = _grads(b, K, d, a, fuse, tables, False, None)
    o2, g2, c2 = _grads(b, K, d, a, fuse, tables, False, 300 * 4 * K * d)        # ~300 nodes (about 13 graphs) per chunk
    assert c1 is None and c2 is not None and len(c2) - 1 >= 5 and c2[0] == 0 and c2[-1] == b["num_nodes"]
    assert torch.equal(o1, o2)
    for name in g1:
        if name in ("dX", "dP"):
            assert torch.equal(g1[name], g2[name]), name
        else:
            assert rel_err(g2[name], g1[name]) < RTOL, (name, rel_err(g2[name], g1[name]))


def test_chunk_boundaries_are_graph_boundaries(lib):
    from kpgnn_b200 import ops
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    gs = synth.zinc_like_graphs(50, seed=5)
    b = collate(gs, (4, 50, 6, 3, 50, 50, "spd"))
    plan, _ = get_plan(b["edge_index"].to(dev), b["edge_attr"].to(dev), b["num_nodes"])
    old = (ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES)
    ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES = 0, 100 * 4 * 4 * 32
    try:
        ch = ops.backward_chunks(plan, 4, 32)
    finally:
        ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES = old
    ends = set()
    tot = 0
    for g in gs:
        tot += g["num_nodes"]
        ends.add(tot)
    assert ch[0] == 0 and ch[-1] == tot and all(c in ends for c in ch[1:]) and ch == sorted(set(ch))
    assert plan.block_ptr is None, "asking for chunk boundaries must not switch the block-resident kernels on"


def test_unchunkable_call_falls_back(lib):
    """eps (GIN self term) is served by another kernel family: the chunk request is ignored, gradients unchanged."""
    from kpgnn_b200.ops import ACT_GELU
    K, d = 4, 64
    b = collate(synth.zinc_like_graphs(40, seed=2), (K, 50, 6, 3, 50, 50, "spd"))
    o1, g1, _ = _grads(b, K, d, ACT_GELU, False, True, True, None)
    o2, g2, c2 = _grads(b, K, d, ACT_GELU, False, True, True, 200 * 4 * K * d)
    assert c2 is not None
    assert torch.equal(o1, o2)
    for name in g1:
        assert torch.equal(g1[name], g2[name]), name
