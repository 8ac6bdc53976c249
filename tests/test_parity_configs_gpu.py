"""GPU parity at the sizes, table shapes and launch geometries the BASELINE.json configs (and the roofline figure)
actually use -- the cases round 1's few-graph layer tests did not reach:
  * the persistent multi-node-per-lane-group paths of the lean kernels (forced with kp_agg_set_launch_geometry on a
    128-graph batch, and at the real 8 192-graph roofline launch against a chunked float64 oracle);
  * configs[3] SR25-shape: 12-regular 25-node graphs, gd kernel, K=4, max_pe_num=1000 -> 1 002-row hop-k tables,
    attrs in the hundreds; widths that force TAB_GLOBAL and the > 200 KB table-gradient fallback;
  * configs[4]: n = 1 280 3-regular graph, K=6 (177 in-entries per node: every node is past the 64-entry window);
  * configs[0] EXP-shape KP-GIN K=3 H=48 on 128 graphs.
Oracle and product run on the same device; bar 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

from kpgnn_b200 import synth
from oracle import layers_torch as OL
from tests.util import RTOL, collate, rel_err, zinc_batch

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
NOISE_ONLY = ("mlp.0.bias", "mlp.3.bias")


def _agg_pair(b, K, d, fuse, dt_oracle=torch.float64, rows0=5, rowsk=52, seed=1, P_on=True):
    """Forward + every gradient of the raw operator (GELU + P + optional fused theta-combine) vs the dense oracle."""
    from kpgnn_b200.ops import khop_aggregate, ACT_GELU
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    N = b["num_nodes"]
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    g = torch.Generator().manual_seed(seed)
    t0 = torch.randn(rows0, d, generator=g).to(dev)
    tk = torch.randn(rowsk, d, generator=g).to(dev)
    th0 = torch.softmax(torch.randn(K, d, generator=g), 0).to(dev)
    x0 = torch.randn(N, K, d, generator=g).to(dev)
    P0 = torch.randn(N, K, d, generator=g).to(dev)
    gy = torch.randn((N, d) if fuse else (N, K, d), generator=g).to(dev)
    outs = []
    for mode in ("oracle", "mine"):
        dt = dt_oracle if mode == "oracle" else torch.float32
        x, P = x0.to(dt).requires_grad_(True), P0.to(dt).requires_grad_(True)
        T0, Tk, th = (t.to(dt).clone().requires_grad_(True) for t in (t0, tk, th0))
        if mode == "oracle":
            z = torch.nn.functional.gelu(OL.dense_khop_aggregate(x, ei, ea, T0, Tk)) + (P if P_on else 0)
            y = (z * th).sum(1) if fuse else z
        else:
            plan, k = get_plan(ei, ea, N)
            y = khop_aggregate(x, plan, k, P=P if P_on else None, T0=T0, Tk=Tk, theta=th if fuse else None,
                               act=ACT_GELU, fuse=fuse)
        y.backward(gy.to(dt))
        outs.append([y.detach(), x.grad, P.grad if P_on else None, T0.grad, Tk.grad] + ([th.grad] if fuse else []))
        del y
    names = ("out", "dX", "dP", "dT0", "dTk", "dtheta")
    for n, a, c in zip(names, outs[1], outs[0]):
        if a is None:
            continue
        assert rel_err(a, c) < RTOL, (n, rel_err(a, c))


@pytest.mark.parametrize("max_ctas,threads", [(8, 1024), (37, 256), (3, 512), (148, 1024)])
@pytest.mark.parametrize("fuse", [True, False])
def test_persistent_multinode_geometry(lib, max_ctas, threads, fuse):
    """The launch geometry of the 8 192-graph roofline call (1024-thread persistent CTAs whose lane groups walk
    MANY nodes through the cross-node software pipeline; B1's per-group dtheta accumulators summed per CTA and
    again across CTAs; B2 as the lean gather) forced onto the 128-graph bench batch: with 8 CTAs x 32 groups every
    lane group processes ~12 nodes."""
    assert lib.kp_agg_set_launch_geometry(max_ctas, threads) == 0
    try:
        _agg_pair(zinc_batch(128, 8, "spd", seed=0), 8, 104, fuse)
    finally:
        lib.kp_agg_set_launch_geometry(0, 0)


@pytest.mark.parametrize("family", [12, 1, 16])
def test_persistent_multinode_other_kernel_families(lib, family):
    """Same forced geometry through the float4 fast kernels, the generic kernels and the TMA-staged forward."""
    assert lib.kp_agg_set_launch_geometry(5, 0) == 0
    lib.kp_agg_set_force_generic(family)
    try:
        _agg_pair(zinc_batch(48, 8, "spd", seed=2), 8, 104, True)
    finally:
        lib.kp_agg_set_force_generic(0)
        lib.kp_agg_set_launch_geometry(0, 0)


def _chunk_bounds(batch_vec, ei, graphs_per_chunk, G):
    """Node / edge ranges of consecutive graph chunks (edges are sorted by graph)."""
    node_ptr = torch.searchsorted(batch_vec, torch.arange(0, G + 1, graphs_per_chunk, device=batch_vec.device).clamp(max=G))
    node_ptr[-1] = batch_vec.numel()
    edge_ptr = torch.searchsorted(ei[0].contiguous(), node_ptr)
    return node_ptr.tolist(), edge_ptr.tolist()


def test_roofline_launch_parity(lib):
    """THE launch bench.py quotes the roofline on: 8 192 ZINC-shaped graphs (189 k nodes, 3.8 M entries), k = 8,
    d = 104, GELU + P + fused geometric combine, production geometry -- forward and the whole backward (dX, dP, dT0,
    dTk, dtheta) against the dense float64 oracle evaluated graph-chunk by graph-chunk (graphs are independent; the
    parameter gradients are summed over chunks in float64)."""
    from kpgnn_b200.data_utils import extract_batch
    from kpgnn_b200.ops import khop_aggregate, ACT_GELU
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    K, d, G = 8, 104, 8192
    graphs = synth.zinc_like_graphs(G, seed=1000 + G)                 # the bench's own roofline batch
    b = extract_batch(graphs, (K, 50, 6, 3, 50, 50, "spd"), dev)      # (extraction is bit-exact-tested elsewhere)
    N = b.num_nodes
    ei, ea = b.edge_index, b.edge_attr
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(N, K, d, device=dev, generator=g).requires_grad_(True)
    P = torch.randn(N, K, d, device=dev, generator=g).requires_grad_(True)
    T0 = torch.randn(5, d, device=dev, generator=g).requires_grad_(True)
    Tk = torch.randn(52, d, device=dev, generator=g).requires_grad_(True)
    th = torch.softmax(torch.randn(K, d, device=dev, generator=g), 0).requires_grad_(True)
    gy = torch.randn(N, d, device=dev, generator=g)
    plan, k = get_plan(ei, ea, N)
    y = khop_aggregate(x, plan, k, P=P, T0=T0, Tk=Tk, theta=th, act=ACT_GELU, fuse=True)
    y.backward(gy)
    node_ptr, edge_ptr = _chunk_bounds(b.batch, ei, 256, G)
    acc = [torch.zeros_like(t, dtype=torch.float64) for t in (T0, Tk, th)]
    worst = {"out": 0.0, "dX": 0.0, "dP": 0.0}
    scale = {"out": float(y.abs().max()), "dX": float(x.grad.abs().max()), "dP": float(P.grad.abs().max())}
    for c in range(len(node_ptr) - 1):
        n0, n1, e0, e1 = node_ptr[c], node_ptr[c + 1], edge_ptr[c], edge_ptr[c + 1]
        xc = x.detach()[n0:n1].double().requires_grad_(True)
        Pc = P.detach()[n0:n1].double().requires_grad_(True)
        t0c, tkc, thc = (t.detach().double().requires_grad_(True) for t in (T0, Tk, th))
        z = torch.nn.functional.gelu(OL.dense_khop_aggregate(xc, ei[:, e0:e1] - n0, ea[e0:e1], t0c, tkc)) + Pc
        yc = (z * thc).sum(1)
        yc.backward(gy[n0:n1].double())
        for name, mine, ref in (("out", y.detach()[n0:n1], yc.detach()), ("dX", x.grad[n0:n1], xc.grad),
                                ("dP", P.grad[n0:n1], Pc.grad)):
            worst[name] = max(worst[name], float((mine.double() - ref).abs().max()) / scale[name])
        for a, t in zip(acc, (t0c, tkc, thc)):
            a += t.grad
        del z, yc
    for name, e in worst.items():
        assert e < RTOL, (name, e)
    # parameter gradients are sums of ~N*k (up to 1.5 M) fp32 terms per element
    for name, mine, ref in (("dT0", T0.grad, acc[0]), ("dTk", Tk.grad, acc[1]), ("dtheta", th.grad, acc[2])):
        assert rel_err(mine, ref) < RTOL, (name, rel_err(mine, ref))


# ------------------------------------------------------------------------------------------------------------------
# configs[3]: SR25-shape
# ------------------------------------------------------------------------------------------------------------------
def _latin_square_srg25():
    """L3(5): a strongly regular graph srg(25,12,5,6) -- the parameter set of the reference's sr25 dataset -- cells of
    a 5x5 grid, adjacent when they share a row, a column or a symbol of the cyclic Latin square."""
    e = [(a, b) for a in range(25) for b in range(25) if a != b and (
        a // 5 == b // 5 or a % 5 == b % 5 or (a // 5 + a % 5) % 5 == (b // 5 + b % 5) % 5)]
    e = np.array(sorted(e), dtype=np.int64).T
    return {"num_nodes": 25, "x": np.ones(25, dtype=np.int64), "edge_index": e, "edge_attr": None, "y": 0.0}


def _sr25_batch(copies=6):
    import networkx as nx
    rng = np.random.default_rng(3)
    base = _latin_square_srg25()
    assert base["edge_index"].shape[1] == 25 * 12
    graphs = [base]
    for s in range(copies - 1):
        if s % 2:
            perm = rng.permutation(25)
            e = perm[base["edge_index"]]
        else:
            G = nx.random_regular_graph(12, 25, seed=s)
            e = np.array(list(G.to_directed().edges)).T
        e = e[:, np.lexsort((e[1], e[0]))]
        graphs.append({"num_nodes": 25, "x": np.ones(25, dtype=np.int64), "edge_index": e.astype(np.int64),
                       "edge_attr": None, "y": 0.0})
    return collate(graphs, (4, 1000, 4, 1, 1000, 1000, "gd"))          # train_SR.py:115-125


def _layer_pair(mine, ora, b, x_shape, P_shape, tol=RTOL):
    dev = torch.device("cuda:0")
    ora.load_state_dict(mine.state_dict())
    mine, ora = mine.to(dev).train(), ora.to(dev).train()
    g = torch.Generator().manual_seed(7)
    N = b["num_nodes"]
    x0 = torch.randn(*[N if s == "N" else s for s in x_shape], generator=g)
    P0 = torch.randn(*[N if s == "N" else s for s in P_shape], generator=g)
    ei, ea, pe = b["edge_index"].to(dev), b["edge_attr"].to(dev), b["pe_attr"]
    # attention combine: oracle in float64 (its nn.LSTM would otherwise be cuDNN's fp32 approximation)
    f64 = any("attention_lstm" in n for n, _ in ora.named_parameters())
    if f64:
        ora = ora.double()
    outs = []
    for layer in (ora, mine):
        dt = torch.float64 if (f64 and layer is ora) else torch.float32
        x, P = x0.clone().to(dev, dt).requires_grad_(True), P0.clone().to(dev, dt).requires_grad_(True)
        y = layer(x * 1.0, ei, ea, pe.to(dev) if pe is not None else None, P)
        y.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(11)).to(dev, dt))
        grads = {"x": x.grad, "P": P.grad}
        grads.update({n: p.grad for n, p in layer.named_parameters()})
        outs.append((y, grads))
    (y0, g0), (y1, g1) = outs
    assert rel_err(y1, y0) < tol, ("forward", rel_err(y1, y0))
    gmax = max(float(v.abs().max()) for v in g0.values() if v is not None)
    for n in g0:
        if g0[n] is None or g1[n] is None:
            for t in (g0[n], g1[n]):
                assert t is None or float(t.abs().max()) == 0.0, n
            continue
        if n in NOISE_ONLY:
            assert float((g1[n] - g0[n]).abs().max()) < 1e-4 * gmax, n
            continue
        err = rel_err(g1[n], g0[n], floor=1e-3 * gmax)
        assert err < tol, (n, err)


@pytest.mark.parametrize("comb", ["geometric", "attention"])
@pytest.mark.parametrize("model", ["KPGCN", "KPGraphSAGE", "KPGIN"])
def test_sr25_shape_layers(lib, model, comb):
    """configs[3] as train_SR.py builds it: hidden 48, K 4, num_hop1_edge 1, max_pe_num 1000 (1 002-row tables, gd
    walk counts in the hundreds as hop-k attrs), aggr add."""
    from kpgnn_b200.layers.KPGCN import KPGCNConv
    from kpgnn_b200.layers.KPGraphSAGE import KPGraphSAGEConv
    from kpgnn_b200.layers.KPGIN import KPGINConv
    torch.manual_seed(0)
    b = _sr25_batch()
    assert int(b["edge_attr"][:, 1:].max()) > 200          # large embedding rows really occur
    H, K = 48, 4
    if model == "KPGCN":
        pair = (KPGCNConv(H, H, K, 1, 1000, comb), OL.OracleKPGCNConv(H, H, K, 1, 1000, comb))
    elif model == "KPGraphSAGE":
        pair = (KPGraphSAGEConv(H, H, K, "add", 1, 1000, comb), OL.OracleKPGraphSAGEConv(H, H, K, "add", 1, 1000, comb))
    else:
        pair = (KPGINConv(H, H, K, 0.1, True, 1, 1000, comb), OL.OracleKPGINConv(H, H, K, 0.1, True, 1, 1000, comb))
    _layer_pair(pair[0], pair[1], b, ("N", H), ("N", K, H // K))


@pytest.mark.parametrize("H", [48, 104])
def test_sr25_shape_large_tables(lib, H):
    """KP-GIN+ on the SR25-shape batch: (3 + 1 002) x H fp32 tables are 193 KB at H = 48 (beyond the 56 KB
    shared-memory staging budget: TAB_GLOBAL lookups forward, sub-table kernel backward) and 418 KB at H = 104
    (beyond 200 KB: the documented float-atomic table-gradient fallback)."""
    from kpgnn_b200.layers.KPGINplus import KPGINPlusConv
    torch.manual_seed(0)
    b = _sr25_batch()
    K = 4
    _layer_pair(KPGINPlusConv(H, H, K, 1, 1000, "geometric"), OL.OracleKPGINPlusConv(H, H, K, 1, 1000, "geometric"), b,
                ("N", K, H), ("N", K, H))


def test_sr25_shape_large_tables_raw_operator(lib):
    _agg_pair(_sr25_batch(), 4, 48, True, rows0=3, rowsk=1002, seed=5)
    _agg_pair(_sr25_batch(), 4, 104, False, rows0=3, rowsk=1002, seed=6)


# ------------------------------------------------------------------------------------------------------------------
# configs[4]: n = 1280 regular graph, K = 6
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def regular1280():
    return collate([synth.regular_graph(1280, 3, 0)], (6, 10, 1, 1, 1, 1, "spd"))      # run_simulation.py:103


def test_regular1280_kgin(lib, regular1280):
    """KGINConv(16, K=6) forward AND backward at n = 1 280 (the reference runs it forward only)."""
    from kpgnn_b200.simulation import KGINConv
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    b = regular1280
    cnt = torch.bincount(b["edge_index"][1], weights=(b["edge_attr"] != 0).sum(1).double(), minlength=1280)
    assert int(cnt.min()) > 64                              # every node is past the lean kernels' entry window
    mine, ora = KGINConv(16, 6, 0.1, True), OL.OracleKGINConv(16, 6, 0.1, True)
    ora.load_state_dict(mine.state_dict())
    mine, ora = mine.to(dev), ora.to(dev)
    ei, ea, bt = b["edge_index"].to(dev), b["edge_attr"].to(dev), b["batch"].to(dev)
    res = []
    for layer in (ora, mine):
        x = torch.ones(b["num_nodes"], 1, device=dev)
        y = layer(x, ei, ea, bt)
        y.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).to(dev))
        res.append((y.detach(), {n: p.grad for n, p in layer.named_parameters()}))
    assert rel_err(res[1][0], res[0][0]) < RTOL
    for n in res[0][1]:
        assert rel_err(res[1][1][n], res[0][1][n]) < RTOL, n


@pytest.mark.parametrize("comb", ["geometric", "attention"])
def test_regular1280_kpgin(lib, regular1280, comb):
    """KP-GIN layer (hidden 96 -> dk 16, K = 6) on the n = 1 280 graph with random features: long rows with tables."""
    from kpgnn_b200.layers.KPGIN import KPGINConv
    torch.manual_seed(0)
    H, K = 96, 6
    _layer_pair(KPGINConv(H, H, K, 0.1, True, 1, 10, comb), OL.OracleKPGINConv(H, H, K, 0.1, True, 1, 10, comb),
                regular1280, ("N", H), ("N", K, H // K))


def test_regular1280_raw_operator_wide(lib, regular1280):
    """Lean-kernel eligible width (d = 104, G = 32) on 177-entry rows: the entry-by-entry path for every node."""
    _agg_pair(regular1280, 6, 104, True, rows0=3, rowsk=12, seed=8)
    _agg_pair(regular1280, 6, 64, False, rows0=3, rowsk=12, seed=9)


# ------------------------------------------------------------------------------------------------------------------
# configs[0]: EXP-shape
# ------------------------------------------------------------------------------------------------------------------
def _exp_like_graphs(num, seed):
    """EXP-shaped inputs (planar SAT-instance graphs: 30-70 nodes, average degree ~2.5, a few dense hubs) when the
    reference's own pickle is not staged; with oracle/_ref present the real first `num` EXP graphs are used."""
    from tests import ref_util as RU
    if RU.available():
        return RU.exp_graphs(num)
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(num):
        n = int(rng.integers(30, 70))
        g = synth.random_typed_graph(rng, n, 2.6 / n, typed=False)
        g["x"] = rng.integers(0, 2, size=n).astype(np.int64)
        out.append(g)
    return out


@pytest.mark.parametrize("comb", ["geometric", "attention"])
def test_exp_shape_kpgin_128_graphs(lib, comb):
    """configs[0]: KP-GIN K=3 H=48 (train_EXP.py defaults: num_hop1_edge 1, max_pe_num 1, spd) on 128 graphs."""
    from kpgnn_b200.layers.KPGIN import KPGINConv
    torch.manual_seed(0)
    b = collate(_exp_like_graphs(128, 4), (3, 1, 5, 1, 1000, 1000, "spd"))          # train_EXP.py:148-158
    H, K = 48, 3
    _layer_pair(KPGINConv(H, H, K, 0., False, 1, 1, comb), OL.OracleKPGINConv(H, H, K, 0., False, 1, 1, comb), b,
                ("N", H), ("N", K, H // K))
