"""The fused KP-GIN+ stack (kpgnn_b200/stack.py: layer-history buffer, strided aggregation inputs, in-place gradient
accumulation, kp_peripheral_grad) against the same backbone run layer by layer through autograd: score, loss, every
parameter gradient, BatchNorm running statistics.  The layer-by-layer path is itself pinned to the reference by
tests/test_golden_gpu.py::test_model_matches_reference_golden (which now also runs through the stack)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def _batch(num_graphs, K, seed, dev):
    from kpgnn_b200.model import Batch
    from tests.util import zinc_batch
    b = zinc_batch(num_graphs, K, "spd", seed=seed)
    b["y"] = torch.randn(num_graphs)
    n = b.pop("num_nodes")
    return Batch(num_nodes=n, **b).to(dev)


@pytest.mark.parametrize("cfg", [(8, 8, "concat", True), (3, 5, "concat", True), (4, 4, "last", True),
                                 (2, 3, "concat", False)])
def test_stack_matches_layerwise(lib, cfg):
    from kpgnn_b200.model import KPGNNPlusRegressor, l1_loss
    K, L, JK, residual = cfg
    dev = torch.device("cuda:0")
    torch.manual_seed(K * 10 + L)
    kw = dict(num_layer=L, hidden_size=104, K=K, input_size=21, num_hop1_edge=3, max_pe_num=50, max_edge_count=50,
              max_hop_num=6, max_distance_count=50, combine="geometric", JK=JK, residual=residual, drop_prob=0.0)
    a = KPGNNPlusRegressor(**kw).to(dev).train()
    with torch.no_grad():                                    # non-trivial BatchNorm affines and combine weights
        for n, p in a.named_parameters():
            if n.endswith("alphas"):
                p.copy_(torch.randn_like(p) * 0.5)
            elif ".mlp.1." in n or ".mlp.4." in n or "norms" in n:
                p.add_(torch.randn_like(p) * 0.2)
    b = copy.deepcopy(a)
    b.embedding_model.use_stack = False
    data = _batch(24, K, 7, dev)
    outs = []
    for m in (a, b):
        for _ in range(2):                                   # running statistics accumulate over two steps
            for p in m.parameters():
                p.grad = None
            score = m(data)
            loss = l1_loss(score, data.y)
            loss.backward()
        outs.append((score.detach(), loss.detach(), {n: p.grad for n, p in m.named_parameters()},
                     {n: v.clone() for n, v in m.named_buffers()}))
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 1e-5 * max(float(outs[1][0].abs().max()), 1.0)
    gmax = max(float(g.abs().max()) for g in outs[1][2].values() if g is not None)
    for n, g in outs[1][2].items():
        h = outs[0][2][n]
        if g is None:
            assert h is None or float(h.abs().max()) == 0.0, n
            continue
        assert h is not None, n
        scale = max(float(g.abs().max()), 1e-3 * gmax)
        assert float((g - h).abs().max()) / scale < 5e-5, (n, float((g - h).abs().max()), scale)
    for n, v in outs[1][3].items():
        w = outs[0][3][n]
        assert float((v.float() - w.float()).abs().max()) <= 1e-5 * max(float(v.float().abs().max()), 1.0), n


def test_stack_deterministic(lib):
    """Bitwise reproducible through the whole backbone (the graph pooling after it is torch's index_add_, which is
    not, and is outside the path)."""
    from kpgnn_b200.model import zinc_kpginplus
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    m = zinc_kpginplus(K=4, num_layer=4).to(dev).train()
    data = _batch(16, 4, 3, dev)
    res = []
    for _ in range(2):
        mm = copy.deepcopy(m)
        mm.embedding_model(data).square().mean().backward()
        res.append([p.grad.clone() for p in mm.parameters() if p.grad is not None])
    for x, y in zip(*res):
        assert torch.equal(x, y)


def test_stack_on_generic_kernels(lib):
    """With the generic kernel family forced, the strided / accumulated dX is refused (rc 3) and the stack falls back
    to a temporary + add; results still match the layer-by-layer path."""
    lib.kp_agg_set_force_generic(1)
    try:
        test_stack_matches_layerwise(lib, (3, 4, "concat", True))
    finally:
        lib.kp_agg_set_force_generic(0)
