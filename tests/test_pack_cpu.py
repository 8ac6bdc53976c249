"""Host re-pack of raw edge lists into the batch CSR (kpgnn_b200.data_utils.pack_csr) against a per-graph dictionary
restatement of the reference's COO -> dense merge (data_utils.py:46-53: duplicate (src, dst) pairs add up, missing
edge_attr means type 2 everywhere).  Pure host logic: runs without a GPU."""
import numpy as np
import pytest

from kpgnn_b200 import data_utils as DU


def naive_pack(graphs):
    rows, off = [], 0
    ns = []
    for g in graphs:
        n, ei, ea = g["num_nodes"], np.asarray(g["edge_index"]).reshape(2, -1), g.get("edge_attr")
        pairs = {}
        for e in range(ei.shape[1]):
            t = 2 if ea is None else int(np.asarray(ea).reshape(-1)[e])
            m, s = pairs.get((int(ei[0, e]), int(ei[1, e])), (0, 0))
            pairs[(int(ei[0, e]), int(ei[1, e]))] = (m + 1, s + t)
        for (a, b), (m, s) in pairs.items():
            rows.append((a + off, b + off, m, s))
        ns.append(n)
        off += n
    rows.sort()
    N = off
    erow = np.zeros(N + 1, dtype=np.int64)
    for a, _, _, _ in rows:
        erow[a + 1] += 1
    return {"N": N, "G": len(ns), "erow": np.cumsum(erow), "ecol": np.array([r[1] for r in rows], dtype=np.int64),
            "emult": np.array([r[2] for r in rows], dtype=np.int64), "etype": np.array([r[3] for r in rows], dtype=np.int64),
            "gptr": np.concatenate([[0], np.cumsum(ns)]), "node_graph": np.repeat(np.arange(len(ns)), ns),
            "pair_off": np.concatenate([[0], np.cumsum(np.array(ns, dtype=np.int64) ** 2)]), "n_max": max(ns) if ns else 0}


def random_graphs(seed, count):
    rng = np.random.default_rng(seed)
    gs = []
    for i in range(count):
        n = int(rng.integers(1, 12))
        e = int(rng.integers(0, 30)) if i % 5 else 0
        ei = rng.integers(0, n, size=(2, e))
        g = {"num_nodes": n, "edge_index": ei}
        if i % 3:
            g["edge_attr"] = rng.integers(0, 5, size=e)
        if i % 4 == 0 and e:
            g["edge_index"] = np.concatenate([ei, ei[:, : e // 2]], axis=1)          # explicit duplicates
            if "edge_attr" in g:
                g["edge_attr"] = np.concatenate([g["edge_attr"], g["edge_attr"][: e // 2] + 1])
        gs.append(g)
    return gs


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pack_matches_per_graph_merge(seed):
    gs = random_graphs(seed, 40)
    got, want = DU.pack_csr(gs), naive_pack(gs)
    assert got["N"] == want["N"] and got["G"] == want["G"] and got["n_max"] == want["n_max"]
    for k in ("erow", "ecol", "emult", "etype", "gptr", "node_graph", "pair_off"):
        assert np.array_equal(np.asarray(got[k], dtype=np.int64), want[k]), k
    assert got["total_pairs"] == int(want["pair_off"][-1])
    assert got["max_type_value"] == (int(want["etype"].max()) if want["etype"].size else 0)


def test_pack_edge_cases():
    empty = DU.pack_csr([])
    assert empty["G"] == 0 and empty["N"] == 0 and empty["erow"].tolist() == [0]
    no_edges = DU.pack_csr([{"num_nodes": 3, "edge_index": np.zeros((2, 0), dtype=np.int64)}] * 2)
    assert no_edges["N"] == 6 and no_edges["erow"].tolist() == [0] * 7 and no_edges["ecol"].size == 0
    with pytest.raises(IndexError):
        DU.pack_csr([{"num_nodes": 2, "edge_index": np.array([[0], [1]])}, {"num_nodes": 2, "edge_index": np.array([[0], [2]])}])
    with pytest.raises(ValueError):
        DU.pack_csr([{"num_nodes": 2, "edge_index": np.array([[0], [1]]), "edge_attr": np.array([1, 2])}])
    with pytest.raises(ValueError):
        DU.pack_csr([{"num_nodes": 2, "edge_index": np.array([[0], [1]]), "edge_attr": np.array([-1])}])


def test_extract_many_splits_a_collated_result_per_graph(monkeypatch):
    """Host logic of kpgnn_b200.data_utils.extract_many (chunking, per-graph node / edge ranges, un-offsetting, the
    edge-less branch) with the device extraction replaced by the numpy oracle: runs without a GPU."""
    import os
    import sys
    import torch
    from kpgnn_b200 import synth
    from oracle.extract_np import extract_multi_hop_neighbors_np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pyg_standin"))
    from torch_geometric.data import Data
    args = (3, 50, 2, 2, 50, 50, "spd")
    graphs = synth.zinc_like_graphs(9, seed=4)
    graphs.insert(3, {"num_nodes": 2, "x": np.zeros(2, dtype=np.int64), "edge_index": np.zeros((2, 0), dtype=np.int64),
                      "edge_attr": np.zeros(0, dtype=np.int64)})
    per_graph = {}

    def fake_upload(csr, device, extra=None):
        return {}

    def fake_extract(csr, K, a1, a2, a3, a4, a5, kern, device=None):
        # collate the oracle's per-graph outputs the way the device path lays them out: node offsets applied, graph-major
        chunk = fake_extract.chunks.pop(0)
        ei, ea, pea, pca, off = [], [], [], [], 0
        for g in chunk:
            o = extract_multi_hop_neighbors_np(g["num_nodes"], g["edge_index"], g["edge_attr"], K, a1, a2, a3, a4, a5, kern)
            ei.append(torch.from_numpy(o["edge_index"]) + off)
            ea.append(torch.from_numpy(o["edge_attr"]))
            pea.append(torch.from_numpy(o["peripheral_edge_attr"]))
            pca.append(torch.from_numpy(o["peripheral_configuration_attr"]))
            off += g["num_nodes"]
        return {"edge_index": torch.cat(ei, 1), "edge_attr": torch.cat(ea), "peripheral_edge_attr": torch.cat(pea),
                "peripheral_configuration_attr": torch.cat(pca), "pe_attr": None, "eptr": None}
    with_edges = [g for g in graphs if g["edge_index"].shape[1]]
    fake_extract.chunks = [with_edges[i:i + 4] for i in range(0, len(with_edges), 4)]
    monkeypatch.setattr(DU, "_upload", fake_upload)
    monkeypatch.setattr(DU, "_extract_device", fake_extract)
    datas = [Data(x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"]),
                  edge_attr=torch.from_numpy(g["edge_attr"])) for g in graphs]
    for d, g in zip(datas, graphs):
        d.num_nodes_ = g["num_nodes"]
    out = DU.extract_many(datas, *args, chunk=4, device="cpu")
    for d, g in zip(out, graphs):
        if g["edge_index"].shape[1] == 0:
            assert d.peripheral_edge_attr.shape == (2, 3, 2, 2) and d.peripheral_configuration.shape == (2, 3, 2)
            continue
        o = extract_multi_hop_neighbors_np(g["num_nodes"], g["edge_index"], g["edge_attr"], *args)
        assert np.array_equal(d.edge_index.numpy(), o["edge_index"]) and np.array_equal(d.edge_attr.numpy(), o["edge_attr"])
        assert np.array_equal(d.peripheral_edge_attr.numpy(), o["peripheral_edge_attr"])
        assert np.array_equal(d.peripheral_configuration_attr.numpy(), o["peripheral_configuration_attr"])
        assert d.pe_attr.shape == (g["num_nodes"], 2) and int(d.pe_attr.abs().sum()) == 0
