"""Closed node blocks of a plan (kp_plan_blocks) and the block-resident kernels built on them (csrc/agg_tile.cu):
bit-exact block boundaries, and forward / backward parity of every layer family that has an unfused aggregation, with
the blocks forced on for short-row molecule batches as well as on the long-row regular graphs they were written for."""
import numpy as np
import pytest
import torch

from kpgnn_b200 import synth
from oracle import layers_torch as OL
from tests.util import RTOL, collate, rel_err, zinc_batch

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def test_plan_blocks_are_the_graphs(lib):
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    gs = synth.zinc_like_graphs(40, seed=3)
    gs.insert(7, {"num_nodes": 3, "x": np.zeros(3, dtype=np.int64), "edge_index": np.zeros((2, 0), dtype=np.int64),
                  "edge_attr": np.zeros(0, dtype=np.int64)})              # three isolated nodes: three blocks
    gs.append(synth.regular_graph(320, 3, 1))
    gs[-1]["edge_attr"] = np.full(gs[-1]["edge_index"].shape[1], 2, dtype=np.int64)
    b = collate(gs, (4, 50, 6, 3, 50, 50, "spd"))
    plan, _ = get_plan(b["edge_index"].to(dev), b["edge_attr"].to(dev), b["num_nodes"])
    bp = plan.blocks().cpu().numpy()[:plan.num_blocks + 1]
    bounds = [0]
    for i, g in enumerate(gs):
        if i == 7:
            bounds += [bounds[-1] + 1, bounds[-1] + 2, bounds[-1] + 3]
        else:
            bounds.append(bounds[-1] + g["num_nodes"])
    assert bp.tolist() == bounds
    assert plan.max_block_nodes == 320
    nnz_last = int(plan.rowptr[-1]) - int(plan.rowptr[bounds[-2] * plan.K])
    assert plan.max_block_nnz == nnz_last


def _pair(b, K, d, act_name, tables, use_P, eps_on, dinv_on=False, seed=0):
    from kpgnn_b200.ops import khop_aggregate, ACT_GELU, ACT_NONE, ACT_RELU
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    N = b["num_nodes"]
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(N, K, d, generator=g).to(dev)
    P0 = torch.randn(N, K, d, generator=g).to(dev)
    t0 = torch.randn(5, d, generator=g).to(dev)
    tk = torch.randn(52, d, generator=g).to(dev)
    gy = torch.randn(N, K, d, generator=g).to(dev)
    eps0 = torch.tensor([0.3], device=dev)
    act = {"gelu": ACT_GELU, "none": ACT_NONE, "relu": ACT_RELU}[act_name]
    f = {"gelu": torch.nn.functional.gelu, "none": lambda z: z, "relu": torch.relu}[act_name]
    res = []
    for mode in ("oracle", "blocks", "rows"):
        x, P = x0.clone().requires_grad_(True), P0.clone().requires_grad_(True)
        T0, Tk, eps = (t.clone().requires_grad_(True) for t in (t0, tk, eps0))
        if mode == "oracle":
            y = f(OL.dense_khop_aggregate(x, ei, ea, T0 if tables else None, Tk if tables else None))
            if use_P:
                y = y + P
            if eps_on:
                y = y + (1 + eps) * x
        else:
            ei2 = ei.clone()                                   # a fresh plan per mode
            plan, k = get_plan(ei2, ea, N)
            if mode == "blocks":
                plan.blocks()
            else:
                plan.block_ptr = None
                import kpgnn_b200.ops as ops
                ops.LONG_ROW_ENTRIES, keep = 10 ** 9, ops.LONG_ROW_ENTRIES
            y = khop_aggregate(x, plan, k, P=P if use_P else None, T0=T0 if tables else None, Tk=Tk if tables else None,
                               eps=eps if eps_on else None, act=act)
            if mode == "rows":
                ops.LONG_ROW_ENTRIES = keep
        y.backward(gy)
        res.append([y.detach(), x.grad] + ([P.grad] if use_P else []) + ([T0.grad, Tk.grad] if tables else []) +
                   ([eps.grad] if eps_on else []))
    for a, c in zip(res[1], res[0]):
        assert rel_err(a, c) < RTOL, rel_err(a, c)
    if not tables:
        # without tables both kernel families add a row's entries in the same order: identical bits forward
        assert torch.equal(res[1][0], res[2][0])
    else:
        assert rel_err(res[1][0], res[2][0]) < 1e-6


@pytest.mark.parametrize("d", [16, 64, 104])
@pytest.mark.parametrize("act,tables,use_P,eps_on", [("none", False, False, True), ("none", True, True, True),
                                                     ("gelu", True, True, False), ("none", True, False, False)])
def test_tile_kernels_on_molecules(lib, d, act, tables, use_P, eps_on):
    _pair(zinc_batch(24, 4, "gd", seed=d), 4, d, act, tables, use_P, eps_on, seed=d)


@pytest.mark.parametrize("d", [16, 48])
def test_tile_kernels_on_regular_graphs(lib, d):
    b = collate([synth.regular_graph(320, 3, s) for s in range(3)], (6, 10, 1, 1, 1, 1, "spd"))
    _pair(b, 6, d, "none", False, False, True, seed=d)
    _pair(b, 6, d, "gelu", True, True, False, seed=d + 1)


@pytest.mark.parametrize("comb", ["geometric", "attention"])
def test_tile_kernels_kpgcn_layer(lib, comb):
    """KP-GCN on a long-row batch: symmetric normalisation (per-entry and per-row dinv) inside the block kernels."""
    from kpgnn_b200.layers.KPGCN import KPGCNConv
    from tests.test_parity_configs_gpu import _layer_pair
    torch.manual_seed(0)
    b = collate([synth.regular_graph(160, 3, s) for s in range(2)], (6, 10, 1, 1, 1, 1, "spd"))
    H, K = 96, 6
    _layer_pair(KPGCNConv(H, H, K, 1, 10, comb), OL.OracleKPGCNConv(H, H, K, 1, 10, comb), b, ("N", H),
                ("N", K, H // K))


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("K,d", [(8, 104), (3, 104), (1, 104), (5, 72), (8, 128)])
def test_block_resident_backward(lib, fuse, K, d):
    """The whole backward as one block-resident kernel (csrc/agg_block_bwd.cu), forced on for a small molecule batch:
    dX, dP, dT0, dTk, dtheta against the dense float64 oracle AND against the three-kernel path (same inputs)."""
    import kpgnn_b200.ops as ops
    from kpgnn_b200.ops import khop_aggregate, ACT_GELU
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    b = zinc_batch(40, K, "spd", seed=K + d)
    N = b["num_nodes"]
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    g = torch.Generator().manual_seed(K)
    t0 = torch.randn(5, d, generator=g).to(dev)
    tk = torch.randn(52, d, generator=g).to(dev)
    th0 = torch.softmax(torch.randn(K, d, generator=g), 0).to(dev)
    x0 = torch.randn(N, K, d, generator=g).to(dev)
    P0 = torch.randn(N, K, d, generator=g).to(dev)
    gy = torch.randn((N, d) if fuse else (N, K, d), generator=g).to(dev)
    res = {}
    for mode in ("oracle", "block", "three"):
        dt = torch.float64 if mode == "oracle" else torch.float32
        x, P = x0.to(dt).clone().requires_grad_(True), P0.to(dt).clone().requires_grad_(True)
        T0, Tk, th = (t.to(dt).clone().requires_grad_(True) for t in (t0, tk, th0))
        if mode == "oracle":
            z = torch.nn.functional.gelu(OL.dense_khop_aggregate(x, ei, ea, T0, Tk if K > 1 else None)) + P
            y = (z * th).sum(1) if fuse else z
        else:
            plan, k = get_plan(ei.clone(), ea, N)
            if mode == "block":
                plan.blocks()
            y = khop_aggregate(x, plan, k, P=P, T0=T0, Tk=Tk if K > 1 else None, theta=th if fuse else None,
                               act=ACT_GELU, fuse=fuse)
            assert (plan.block_ptr is not None) == (mode == "block")
        y.backward(gy.to(dt))
        res[mode] = [x.grad, P.grad, T0.grad] + ([Tk.grad] if K > 1 else []) + ([th.grad] if fuse else [])
    for a, c in zip(res["block"], res["oracle"]):
        assert rel_err(a, c) < RTOL, rel_err(a, c)
    for a, c in zip(res["block"], res["three"]):
        assert rel_err(a, c) < 1e-6


def test_block_resident_backward_in_stack(lib):
    """The layer-history stack (strided, accumulated dX) over the block-resident backward: same gradients as without."""
    import kpgnn_b200.ops as ops
    from kpgnn_b200.model import Batch, l1_loss, zinc_kpginplus
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = zinc_kpginplus(8, 8, 104).to(dev).train()
    d = zinc_batch(16, 8, "spd", seed=3)
    res = []
    for force in (False, True):
        keep = ops.BLOCK_BWD_MIN_BYTES
        ops.BLOCK_BWD_MIN_BYTES = 0 if force else None
        try:
            b = Batch(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}).to(dev)
            for p in model.parameters():
                p.grad = None
            loss = l1_loss(model(b), b.y)
            loss.backward()
            res.append([loss.detach()] + [p.grad.clone() for n, p in model.named_parameters()
                                          if p.grad is not None and not n.endswith(("mlp.0.bias", "mlp.3.bias"))])
        finally:
            ops.BLOCK_BWD_MIN_BYTES = keep
    gmax = max(float(t.abs().max()) for t in res[0][1:])
    for a, c in zip(res[1], res[0]):
        # (the scalar gates pew / pcw are near-cancelling sums of ~5e5 terms: 1e-4)
        assert rel_err(a, c, floor=1e-3 * gmax) < (1e-4 if a.numel() == 1 else 2e-5)
