"""GPU parity of the CUDA extractor (kp_extract_* through the C ABI).  Integer work: BIT-EXACT against
(a) the reference's own outputs in tests/golden/extract.npz, (b) the numpy oracle on seeded inputs, and
(c) size-independent properties at the BASELINE.json full sizes."""
import sys
import os

import numpy as np
import pytest
import torch

from kpgnn_b200 import synth
from oracle.extract_np import extract_multi_hop_neighbors_np
from tests import golden_util as GU
from tests.util import collate

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pyg_standin"))


def _data(g, device="cpu"):
    from torch_geometric.data import Data          # the test-only stand-in: an attribute bag like PyG's Data
    d = Data(x=torch.from_numpy(g["x"]) if g.get("x") is not None else None,
             edge_index=torch.from_numpy(np.asarray(g["edge_index"])).to(device),
             edge_attr=None if g["edge_attr"] is None else torch.from_numpy(g["edge_attr"]).to(device))
    d.num_nodes_ = g["num_nodes"]
    return d


def test_reference_goldens_bit_exact(lib):
    from kpgnn_b200.data_utils import extract_multi_hop_neighbors
    z, meta = GU.load("extract.npz")
    for i, m in enumerate(meta):
        pre = "c%d_" % i
        g = {"num_nodes": m["num_nodes"], "x": np.zeros(m["num_nodes"], dtype=np.int64),
             "edge_index": z[pre + "in_edge_index"], "edge_attr": z[pre + "in_edge_attr"] if m["typed"] else None}
        d = extract_multi_hop_neighbors(_data(g), *m["args"])
        for f in ("edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr",
                  "peripheral_configuration"):
            got = d._store.get(f, None)
            if f in ("edge_index", "edge_attr") and g["edge_index"].shape[1] == 0:
                continue
            if f in m["fields"]:
                ref = z[pre + "out_" + f]
                assert got is not None, (m["name"], f)
                assert got.dtype == torch.long and tuple(got.shape) == ref.shape, (m["name"], f, got.shape, ref.shape)
                assert np.array_equal(got.cpu().numpy(), ref), (m["name"], f)
            else:
                assert got is None, (m["name"], f)


def _check_batch(graphs, args):
    from kpgnn_b200.data_utils import extract_batch
    b = extract_batch(graphs, args, "cuda:0")
    ref = collate(graphs, args)
    assert b.num_nodes == ref["num_nodes"] and b.num_graphs == ref["num_graphs"]
    for f in ("edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr", "batch"):
        got, exp = getattr(b, f), ref[f]
        if exp is None:
            assert got is None, f
        else:
            assert got is not None and got.dtype == torch.long, f
            assert tuple(got.shape) == tuple(exp.shape), (f, got.shape, exp.shape)
            assert torch.equal(got.cpu(), exp), f
    return b


@pytest.mark.parametrize("args", [(8, 50, 6, 3, 50, 50, "spd"), (16, 50, 6, 3, 50, 50, "spd"),
                                  (4, 50, 6, 3, 50, 50, "gd"), (1, 50, 6, 3, 50, 50, "spd"),
                                  (3, 1, 2, 1, 1, 1, "gd"), (5, 3, 0, 3, 50, 50, "spd")])
def test_zinc_batches_vs_oracle(lib, args):
    _check_batch(synth.zinc_like_graphs(48, seed=args[0]), args)


def test_random_typed_directed_multigraphs_vs_oracle(lib):
    rng = np.random.default_rng(5)
    for it in range(12):
        graphs = []
        for i in range(6):
            g = synth.random_typed_graph(rng, int(rng.integers(1, 40)), float(rng.uniform(0.03, 0.5)),
                                         num_types=int(rng.integers(1, 6)), directed=bool(rng.integers(0, 2)),
                                         typed=True)
            if g["edge_index"].shape[1] and it % 3 == 0:
                # duplicate a few edges and add self loops: the reference sums duplicates (data_utils.py:52-53)
                ei, ea = g["edge_index"], g["edge_attr"]
                dup = rng.integers(0, ei.shape[1], size=3)
                loops = rng.integers(0, g["num_nodes"], size=2)
                g["edge_index"] = np.concatenate([ei, ei[:, dup], np.stack([loops, loops])], 1)
                g["edge_attr"] = np.concatenate([ea, ea[dup], np.full(2, 3)])
            graphs.append(g)
        args = (int(rng.integers(1, 7)), int(rng.choice([1, 3, 50, 1000])), int(rng.integers(1, 5)),
                int(rng.integers(1, 5)), int(rng.choice([1, 3, 50])), int(rng.choice([1, 3, 50])),
                "spd" if it % 2 else "gd")
        _check_batch(graphs, args)


def test_cutoff_one_configuration_vs_oracle(lib):
    """max_hop_num = 1 (run_simulation.py:103) takes the lane-per-member path of extract_peripheral_kernel: typed, directed,
    duplicate and self-loop edges, dense and sparse graphs, both kernels -- bit-exact against the oracle."""
    import networkx as nx
    rng = np.random.default_rng(11)
    for it in range(10):
        graphs = []
        for i in range(6):
            g = synth.random_typed_graph(rng, int(rng.integers(2, 48)), float(rng.uniform(0.05, 0.6)),
                                         num_types=int(rng.integers(1, 6)), directed=bool(rng.integers(0, 2)), typed=True)
            if g["edge_index"].shape[1] and it % 2 == 0:
                ei, ea = g["edge_index"], g["edge_attr"]
                dup = rng.integers(0, ei.shape[1], size=4)
                loops = rng.integers(0, g["num_nodes"], size=3)
                g["edge_index"] = np.concatenate([ei, ei[:, dup], np.stack([loops, loops])], 1)
                g["edge_attr"] = np.concatenate([ea, ea[dup], np.full(3, 2)])
            graphs.append(g)
        args = (int(rng.integers(1, 7)), int(rng.choice([1, 3, 50])), 1, int(rng.integers(1, 5)),
                int(rng.choice([1, 3, 50])), int(rng.choice([1, 3, 50])), "spd" if it % 2 else "gd")
        _check_batch(graphs, args)
    dense = []
    for s in range(3):
        G = nx.random_regular_graph(12, 25, seed=s)
        e = np.array(list(G.to_directed().edges)).T
        e = e[:, np.lexsort((e[1], e[0]))]
        dense.append({"num_nodes": 25, "x": np.ones(25, dtype=np.int64), "edge_index": e.astype(np.int64), "edge_attr": None})
    _check_batch(dense, (4, 1000, 1, 2, 1000, 1000, "spd"))


def test_dense_regular_gd_saturation_vs_oracle(lib):
    """SR25-shape: 25 nodes, 12-regular, gd K=4 -- walk counts reach the hundreds (int16 attrs, cap 1000)."""
    import networkx as nx
    graphs = []
    for s in range(4):
        G = nx.random_regular_graph(12, 25, seed=s)
        e = np.array(list(G.to_directed().edges)).T
        e = e[:, np.lexsort((e[1], e[0]))]
        graphs.append({"num_nodes": 25, "x": np.ones(25, dtype=np.int64), "edge_index": e.astype(np.int64),
                       "edge_attr": None})
    for kern in ("spd", "gd"):
        _check_batch(graphs, (4, 1000, 4, 1, 1000, 1000, kern))          # train_SR.py:115-125
    _check_batch(graphs, (6, 30, 2, 2, 7, 9, "gd"))                     # saturating caps


def test_regular_320_vs_oracle(lib):
    _check_batch([synth.regular_graph(320, 3, s) for s in range(2)], (6, 10, 1, 1, 1, 1, "spd"))


def test_batch_with_edgeless_graph(lib):
    gs = synth.zinc_like_graphs(3, seed=4)
    gs.insert(1, {"num_nodes": 5, "x": np.zeros(5, dtype=np.int64), "edge_index": np.zeros((2, 0), dtype=np.int64),
                  "edge_attr": np.zeros(0, dtype=np.int64)})
    from kpgnn_b200.data_utils import extract_batch
    b = extract_batch(gs, (4, 50, 6, 3, 50, 50, "spd"), "cuda:0")
    n0 = gs[0]["num_nodes"]
    assert int(b.peripheral_edge_attr[n0:n0 + 5].abs().sum()) == 0
    assert not bool(((b.edge_index >= n0) & (b.edge_index < n0 + 5)).any())


def test_full_size_properties(lib):
    """BASELINE.json full sizes (regular n=1280 K=6; 128 ZINC-shaped graphs K=8 and K=16): properties that do
    not need the oracle."""
    from kpgnn_b200.data_utils import extract_batch
    cases = [([synth.regular_graph(1280, 3, 0)], (6, 10, 1, 1, 1, 1, "spd")),
             (synth.zinc_like_graphs(128, seed=0), (8, 50, 6, 3, 50, 50, "spd")),
             (synth.zinc_like_graphs(128, seed=0), (16, 50, 6, 3, 50, 50, "spd"))]
    for graphs, args in cases:
        a = extract_batch(graphs, args, "cuda:0")
        b = extract_batch(graphs, args, "cuda:0")
        ei, ea = a.edge_index.cpu(), a.edge_attr.cpu()
        for f in ("edge_index", "edge_attr", "peripheral_edge_attr", "peripheral_configuration_attr"):
            assert torch.equal(getattr(a, f), getattr(b, f)), f                  # run-to-run identical
        N = a.num_nodes
        key = ei[0] * N + ei[1]
        assert bool((key[1:] > key[:-1]).all())                                  # strictly (src,dst)-sorted
        assert not bool((ei[0] == ei[1]).any())                                  # no self pairs
        assert bool(((ea != 0).sum(1) == 1).all())                               # spd: exactly one hop per pair
        assert bool((a.batch.cpu()[ei[0]] == a.batch.cpu()[ei[1]]).all())        # never crosses graphs
        # undirected inputs: the K-hop relation and its attrs are symmetric
        rev = torch.searchsorted(key, ei[1] * N + ei[0])
        assert torch.equal(key[rev], ei[1] * N + ei[0]) and torch.equal(ea[rev], ea)
        # hop-1 edges are exactly the input edges with their types
        raw_src = np.concatenate([g["edge_index"][0] + o for g, o in zip(graphs, np.cumsum([0] + [g["num_nodes"] for g in graphs[:-1]]))])
        raw_dst = np.concatenate([g["edge_index"][1] + o for g, o in zip(graphs, np.cumsum([0] + [g["num_nodes"] for g in graphs[:-1]]))])
        hop1 = ea[:, 0] != 0
        assert int(hop1.sum()) == raw_src.size
        assert np.array_equal(np.sort(key[hop1].numpy()), np.sort(raw_src * N + raw_dst))
        if args[0] > 1:
            assert int(a.pe_attr.abs().sum()) == 0
        assert int(ea[:, 1:].max()) <= args[1] + 1
        assert int(a.peripheral_configuration_attr.max()) <= args[5]
    # and one oracle spot check at full size for a single molecule batch entry
    g = synth.zinc_like_graphs(128, seed=0)[17]
    ref = extract_multi_hop_neighbors_np(g["num_nodes"], g["edge_index"], g["edge_attr"], 8, 50, 6, 3, 50, 50, "spd")
    one = extract_batch([g], (8, 50, 6, 3, 50, 50, "spd"), "cuda:0")
    assert np.array_equal(one.peripheral_configuration_attr.cpu().numpy(), ref["peripheral_configuration_attr"])


def test_reference_goldens_full_configs_bit_exact(lib):
    """tests/golden/extract_full.npz: the reference's own outputs on the first 128 graphs of its EXP dataset
    (configs[0]), all 15 SR25 graphs under gd and spd (configs[3]) and an n = 1 280 3-regular graph at K = 6
    (configs[4]); every field bit-exact, extracted here as ONE batch per config (graph-major collation)."""
    from kpgnn_b200.data_utils import extract_batch
    z, meta = GU.load("extract_full.npz")
    groups = {}
    for i, m in enumerate(meta):
        groups.setdefault(tuple(m["args"]), []).append(i)
    for args, idxs in groups.items():
        graphs = [{"num_nodes": meta[i]["num_nodes"], "x": np.zeros(meta[i]["num_nodes"], dtype=np.int64),
                   "edge_index": z["c%d_in_edge_index" % i].astype(np.int64), "edge_attr": None} for i in idxs]
        b = extract_batch(graphs, args, "cuda:0")
        off = 0
        exp = {f: [] for f in ("edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr",
                               "peripheral_configuration_attr")}
        for i, g in zip(idxs, graphs):
            for f in exp:
                v = z["c%d_out_%s" % (i, f)].astype(np.int64)
                exp[f].append(v + off if f == "edge_index" else v)
            off += g["num_nodes"]
        for f, parts in exp.items():
            ref = np.concatenate(parts, axis=1 if f == "edge_index" else 0)
            got = getattr(b, f).cpu().numpy()
            assert got.shape == ref.shape and np.array_equal(got, ref), (args, f)


def test_extract_many_equals_per_graph_calls(lib):
    """The batched dataset pre-transform (kpgnn_b200.data_utils.extract_many, what datasets/*.py process() runs over a whole
    data list) gives every graph exactly the fields of the per-graph drop-in -- including graphs without edges, which take
    the reference's early-return branch (data_utils.py:37-44) -- for typed, untyped, directed and duplicate-edge inputs."""
    from kpgnn_b200.data_utils import extract_many, extract_multi_hop_neighbors
    rng = np.random.default_rng(3)
    graphs = synth.zinc_like_graphs(20, seed=9)
    for i in range(8):
        graphs.append(synth.random_typed_graph(rng, int(rng.integers(2, 30)), float(rng.uniform(0.05, 0.5)),
                                               num_types=3, directed=bool(i % 2), typed=bool(i % 3)))
    graphs.insert(5, {"num_nodes": 4, "x": np.zeros(4, dtype=np.int64), "edge_index": np.zeros((2, 0), dtype=np.int64),
                      "edge_attr": np.zeros(0, dtype=np.int64)})
    for args in ((4, 50, 6, 3, 50, 50, "spd"), (3, 10, 1, 1, 1, 1, "gd"), (1, 50, 2, 2, 50, 50, "spd")):
        one = [extract_multi_hop_neighbors(_data(g), *args) for g in graphs]
        many = extract_many([_data(g) for g in graphs], *args, chunk=7)
        for a, b in zip(one, many):
            for f in ("edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr",
                      "peripheral_configuration"):
                va, vb = getattr(a, f, None), getattr(b, f, None)
                assert (va is None) == (vb is None), f
                if va is not None:
                    assert va.dtype == vb.dtype and va.shape == vb.shape and torch.equal(va, vb), f
