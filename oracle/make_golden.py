"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported behind the
torch_geometric stand-in, oracle/refimport.py) in the build container.  The reference cannot travel to the GPU
box, so its outputs are committed as small fixtures together with this script.

    python -m oracle.make_golden            # rewrites tests/golden/

Fixtures:
  extract.npz   extraction inputs/outputs of data_utils.extract_multi_hop_neighbors (bit-exact targets)
  layers.npz    state_dict + inputs + output + gradients of each layers/* module (1e-5 relative targets)
  model_zinc.npz   GraphRegression(GNNPlus(KPGINPlus K=8 L=8 H=104)) on an 8-graph ZINC-shaped batch
  extract_full.npz 128 EXP graphs, the 15 SR25 graphs (gd + spd), one n = 1 280 regular graph: extraction at real inputs
  models_cfg.npz   the other BASELINE.json model configs (EXP KP-GIN, KPGINPrime, SR25 KPGCN / KPGraphSAGE): reference
                   state_dict + batch + prediction + loss + every parameter gradient

    python -m oracle.make_golden [extract] [extract_full] [layers] [model] [models_cfg]     # subset
"""
import argparse
import json
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kpgnn_b200 import synth  # noqa: E402
from oracle import refimport  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _ref_extract(ns, g, args):
    d = ns.Data(x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"]),
                edge_attr=None if g["edge_attr"] is None else torch.from_numpy(g["edge_attr"]))
    d.num_nodes_ = g["num_nodes"]
    return ns.data_utils.extract_multi_hop_neighbors(d, *args)


def extraction_cases():
    rng = np.random.default_rng(2024)
    cases = []
    zg = synth.zinc_like_graphs(4, seed=11)
    for g in zg:
        cases.append(("zinc_spd8", g, (8, 50, 6, 3, 50, 50, "spd")))
    cases.append(("zinc_gd4", zg[0], (4, 50, 6, 3, 50, 50, "gd")))
    cases.append(("zinc_spd16", zg[1], (16, 50, 6, 3, 50, 50, "spd")))
    cases.append(("zinc_k1", zg[2], (1, 50, 6, 3, 50, 50, "spd")))
    cases.append(("zinc_noperiph", zg[3], (3, 50, 0, 3, 50, 50, "spd")))
    cases.append(("zinc_caps1", zg[3], (5, 1, 2, 1, 1, 1, "gd")))
    with open(os.path.join(refimport.REF_ROOT, "data/EXP/raw/GRAPHSAT.pkl"), "rb") as f:
        exp = pickle.load(f)
    for i in (0, 1, 600):
        d = exp[i]
        g = {"num_nodes": d.x.size(0), "x": d.x[:, 0].long().numpy(), "edge_index": d.edge_index.numpy(),
             "edge_attr": None}
        cases.append(("exp%d_spd3" % i, g, (3, 1, 5, 1, 1000, 1000, "spd")))       # train_EXP.py:148-158
        if i == 0:
            cases.append(("exp%d_gd3" % i, g, (3, 1, 5, 1, 1000, 1000, "gd")))
    import networkx as nx
    sr = nx.read_graph6(os.path.join(refimport.REF_ROOT, "data/sr25/raw/sr251256.g6"))
    for i in (0, 7):
        e = np.array(list(sr[i].to_directed().edges)).T
        e = e[:, np.lexsort((e[1], e[0]))]
        g = {"num_nodes": 25, "x": np.ones(25, dtype=np.int64), "edge_index": e.astype(np.int64), "edge_attr": None}
        cases.append(("sr25_%d_spd4" % i, g, (4, 1000, 4, 1, 1000, 1000, "spd")))  # train_SR.py:115-125
        cases.append(("sr25_%d_gd4" % i, g, (4, 1000, 4, 1, 1000, 1000, "gd")))
    cases.append(("regular40", synth.regular_graph(40, 3, 0), (6, 10, 1, 1, 1, 1, "spd")))   # run_simulation.py:103
    cases.append(("regular160", synth.regular_graph(160, 3, 1), (4, 10, 1, 1, 1, 1, "spd")))
    for i in range(6):
        g = synth.random_typed_graph(rng, int(rng.integers(5, 28)), float(rng.uniform(0.08, 0.4)),
                                     num_types=int(rng.integers(1, 5)), directed=bool(i % 2), typed=bool(i % 3))
        if g["edge_index"].shape[1] == 0:
            continue
        cases.append(("gnp%d" % i, g, (int(rng.integers(2, 6)), int(rng.choice([1, 3, 50])), int(rng.integers(1, 4)),
                                      int(rng.integers(1, 4)), int(rng.choice([1, 3, 50])), int(rng.choice([2, 50])),
                                      "spd" if i % 2 else "gd")))
    iso = {"num_nodes": 6, "x": np.zeros(6, dtype=np.int64),
           "edge_index": np.array([[0, 1, 1, 2], [1, 0, 2, 1]], dtype=np.int64), "edge_attr": None}
    cases.append(("isolated_nodes", iso, (3, 5, 2, 2, 5, 5, "spd")))
    empty = {"num_nodes": 4, "x": np.zeros(4, dtype=np.int64), "edge_index": np.zeros((2, 0), dtype=np.int64),
             "edge_attr": None}
    cases.append(("no_edges", empty, (3, 5, 2, 2, 5, 5, "spd")))
    return cases


def make_extract(ns):
    store, meta = {}, []
    for idx, (name, g, args) in enumerate(extraction_cases()):
        r = _ref_extract(ns, g, args)
        pre = "c%d_" % idx
        store[pre + "in_edge_index"] = g["edge_index"]
        if g["edge_attr"] is not None:
            store[pre + "in_edge_attr"] = g["edge_attr"]
        fields = []
        for k in ("edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr",
                  "peripheral_configuration"):
            v = r._store.get(k, None)
            if k in ("edge_index", "edge_attr") and g["edge_index"].shape[1] == 0:
                continue
            if v is not None:
                store[pre + "out_" + k] = v.contiguous().numpy()
                fields.append(k)
        meta.append({"name": name, "num_nodes": int(g["num_nodes"]), "args": list(args), "fields": fields,
                     "typed": g["edge_attr"] is not None})
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "extract.npz"), **store)
    print("extract.npz: %d cases" % len(meta))


def make_extract_full(ns):
    """BASELINE.json configs at their real inputs: the first 128 graphs of the in-repo EXP dataset (configs[0]), all 15
    in-repo SR25 graphs under both kernels (configs[3]) and one n = 1 280 3-regular graph at K = 6 (configs[4]); outputs
    stored in the narrowest integer type that holds them (the npz stays < 1 MB)."""
    from tests import ref_util as RU
    cases = [("exp%d_spd3" % i, g, (3, 1, 5, 1, 1000, 1000, "spd")) for i, g in enumerate(RU.exp_graphs(128))]
    for i, g in enumerate(RU.sr25_graphs()):
        cases.append(("sr25_%d_gd4" % i, g, (4, 1000, 4, 1, 1000, 1000, "gd")))
        cases.append(("sr25_%d_spd4" % i, g, (4, 1000, 4, 1, 1000, 1000, "spd")))
    cases.append(("regular1280_spd6", synth.regular_graph(1280, 3, 0), (6, 10, 1, 1, 1, 1, "spd")))

    def narrow(a):
        a = np.ascontiguousarray(a)
        for dt in (np.uint8, np.uint16, np.int32):
            if a.size == 0 or (a.min() >= np.iinfo(dt).min and a.max() <= np.iinfo(dt).max):
                return a.astype(dt)
        return a
    store, meta = {}, []
    for idx, (name, g, args) in enumerate(cases):
        r = _ref_extract(ns, g, args)
        pre = "c%d_" % idx
        store[pre + "in_edge_index"] = narrow(g["edge_index"])
        fields = []
        for k in ("edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr"):
            v = r._store.get(k, None)
            if v is not None:
                assert v.dtype == torch.long
                store[pre + "out_" + k] = narrow(v.contiguous().numpy())
                fields.append(k)
        meta.append({"name": name, "num_nodes": int(g["num_nodes"]), "args": list(args), "fields": fields})
        if idx % 32 == 0:
            print("  extract_full: %d / %d" % (idx, len(cases)), flush=True)
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "extract_full.npz"), **store)
    print("extract_full.npz: %d cases" % len(meta))


def _collate(ns, graphs, args):
    datas = [_ref_extract(ns, g, args) for g in graphs]
    return ns.Batch.from_data_list(datas)


def make_layers(ns):
    torch.manual_seed(1234)
    store, meta = {}, []
    graphs = synth.zinc_like_graphs(4, seed=5)

    def add(name, layer, K, kern, x_shape, p_shape, ctor, gine=False, pe_random=False):
        b = _collate(ns, graphs, (K, 50, 6, 3, 50, 50, kern))
        N = b.num_nodes
        g = torch.Generator().manual_seed(99)
        x = torch.randn(*[N if s == "N" else s for s in x_shape], generator=g).requires_grad_(True)
        P = None
        if p_shape is not None:
            P = torch.randn(*[N if s == "N" else s for s in p_shape], generator=g).requires_grad_(True)
        pe = b.pe_attr if (K > 1 and not gine) else None
        if pe_random and K > 1:
            pe = torch.randint(0, 5, (N, K - 1), generator=g)
        layer.train()
        if gine:
            y = layer(x * 1.0, b.edge_index, b.edge_attr[:, :1])
        else:
            y = layer(x * 1.0, b.edge_index, b.edge_attr, pe, P)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        pre = "l%d_" % len(meta)
        store[pre + "edge_index"] = b.edge_index.numpy()
        store[pre + "edge_attr"] = b.edge_attr.numpy()
        store[pre + "x"] = x.detach().numpy()
        store[pre + "gy"] = gy.numpy()
        store[pre + "y"] = y.detach().numpy()
        store[pre + "gx"] = x.grad.numpy()
        if P is not None:
            store[pre + "P"] = P.detach().numpy()
            store[pre + "gP"] = P.grad.numpy()
        if pe is not None:
            store[pre + "pe"] = pe.numpy()
        for k, v in layer.state_dict().items():
            store[pre + "sd_" + k] = v.numpy()
        for k, p in layer.named_parameters():
            if p.grad is not None:
                store[pre + "gp_" + k] = p.grad.numpy()
        meta.append({"name": name, "ctor": ctor, "K": K, "gine": gine, "N": N})

    for comb in ("geometric", "attention"):
        add("KPGINConv_" + comb, ns.KPGIN.KPGINConv(32, 32, 4, 0.1, True, 3, 50, comb), 4, "spd", ("N", 32),
            ("N", 4, 8), ["KPGINConv", 32, 32, 4, 0.1, True, 3, 50, comb])
        add("KPGINPlusConv_" + comb, ns.KPGINplus.KPGINPlusConv(24, 24, 4, 3, 50, comb), 4, "gd", ("N", 4, 24),
            ("N", 4, 24), ["KPGINPlusConv", 24, 24, 4, 3, 50, comb])
        add("KPGCNConv_" + comb, ns.KPGCN.KPGCNConv(32, 32, 4, 3, 50, comb), 4, "gd", ("N", 32), ("N", 4, 8),
            ["KPGCNConv", 32, 32, 4, 3, 50, comb])
        add("KPGraphSAGEConv_" + comb, ns.KPGraphSAGE.KPGraphSAGEConv(32, 32, 4, "add", 3, 50, comb), 4, "spd",
            ("N", 32), ("N", 4, 8), ["KPGraphSAGEConv", 32, 32, 4, "add", 3, 50, comb])
    add("KPGINPlusConv_zinc", ns.KPGINplus.KPGINPlusConv(104, 104, 8, 3, 50, "geometric"), 8, "spd", ("N", 8, 104),
        ("N", 8, 104), ["KPGINPlusConv", 104, 104, 8, 3, 50, "geometric"])
    add("KPGINConv_prime", ns.KPGIN.KPGINConv(96, 96, 16, 0., False, 3, 50, "geometric"), 16, "spd", ("N", 96),
        ("N", 16, 6), ["KPGINConv", 96, 96, 16, 0., False, 3, 50, "geometric"])
    add("KPGINConv_pe", ns.KPGIN.KPGINConv(32, 32, 4, 0.0, False, 3, 50, "geometric"), 4, "spd", ("N", 32),
        ("N", 4, 8), ["KPGINConv", 32, 32, 4, 0.0, False, 3, 50, "geometric"], pe_random=True)
    add("KPGINPlusConv_k1", ns.KPGINplus.KPGINPlusConv(24, 24, 1, 3, 50, "geometric"), 1, "spd", ("N", 1, 24),
        ("N", 1, 24), ["KPGINPlusConv", 24, 24, 1, 3, 50, "geometric"])
    add("GINEConv", ns.gine.GINEConv(40, 40, 0.2, 3, True), 8, "spd", ("N", 40), None,
        ["GINEConv", 40, 40, 0.2, 3, True], gine=True)
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "layers.npz"), **store)
    print("layers.npz: %d cases" % len(meta))


def make_model(ns):
    torch.manual_seed(4321)
    args = argparse.Namespace(model_name="KPGINPlus", hidden_size=104, K=8, num_hop1_edge=3, max_pe_num=50,
                              combine="geometric", num_layer=8, eps=0., train_eps=False, aggr="add")
    layer = ns.layer_utils.make_gnn_layer(args)
    gnn = ns.GNNs.GNNPlus(num_layer=8, gnn_layer=layer, JK="concat", norm_type="Batch",
                          init_emb=ns.input_encoder.EmbeddingEncoder(21, 104), residual=True, virtual_node=False,
                          use_rd=False, num_hop1_edge=3, max_edge_count=50, max_hop_num=6, max_distance_count=50,
                          wo_peripheral_edge=False, wo_peripheral_configuration=False, drop_prob=0.0)
    model = ns.GraphRegression.GraphRegression(embedding_model=gnn, pooling_method="sum")
    model.reset_parameters()
    # move the model off its all-zero alphas / fresh BN so every gradient path is exercised
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("alphas"):
                p.add_(0.3 * torch.randn_like(p))
    graphs = synth.zinc_like_graphs(8, seed=21)
    b = _collate(ns, graphs, (8, 50, 6, 3, 50, 50, "spd"))
    b.y = torch.tensor([g["y"] for g in graphs], dtype=torch.float32)
    model.train()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    score = model(b)
    loss = (score.squeeze() - b.y.squeeze()).abs().mean()
    loss.backward()
    store = {"score": score.detach().numpy(), "loss": np.array(loss.item(), dtype=np.float32)}
    for k in ("x", "edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr",
              "batch", "y"):
        store["b_" + k] = b._store[k].contiguous().numpy()
    for k, v in sd.items():
        store["sd_" + k] = v.numpy()
    for k, p in model.named_parameters():
        store["gp_" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "model_zinc.npz"), **store)
    print("model_zinc.npz: loss %.6f, %d params" % (loss.item(), sum(p.numel() for p in model.parameters())))


def make_models_cfg(ns):
    """BASELINE.json configs[0], [2], [3] as the reference's train scripts build them (tests/ref_util.py CONFIGS),
    forward + backward on the CPU, for the GPU test that loads the same state_dict into the product.  The stored
    truth is the reference evaluated in FLOAT64; next to every tensor goes `own32`, the distance of the reference's own
    fp32 evaluation from it (deep BatchNorm stacks amplify fp32 rounding -- a property of the model; the GPU test
    holds the product to max(1e-5, 10 x own32))."""
    import copy
    from tests import ref_util as RU

    def rel(a, b, floor):
        return float((a.double() - b.double()).abs().max()) / max(float(b.abs().max()), floor, 1e-30)
    store, meta = {}, []
    for name, graphs in (("exp", RU.exp_graphs(16)), ("prime", synth.zinc_like_graphs(6, seed=41)),
                         ("sr_gcn", RU.sr25_graphs(random_x=True)[:6]), ("sr_sage", RU.sr25_graphs(random_x=True)[:6])):
        cfg = RU.CONFIGS[name]
        torch.manual_seed(99)
        b = RU.ref_batch(ns, graphs, cfg["extract"], torch.float32 if cfg["head"][0] == "regression" else torch.int64)
        model = RU.build_model(cfg, ns.GNNs, ns.layer_utils.make_gnn_layer, ns.input_encoder.EmbeddingEncoder, ns)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("alphas"):
                    p.add_(0.3 * torch.randn_like(p))
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        res = {}
        for dt in (torch.float32, torch.float64):
            m = copy.deepcopy(model).to(dt).train()
            bb = b.clone()
            if bb.y.dtype == torch.float32:
                bb.y = bb.y.to(dt)
            pred = m(bb)
            loss = RU.loss_fn(cfg, pred, bb.y)
            loss.backward()
            res[dt] = (pred.detach(), float(loss), {k: p.grad for k, p in m.named_parameters() if p.grad is not None})
        p64, l64, g64 = res[torch.float64]
        p32, l32, g32 = res[torch.float32]
        gmax = max(float(v.abs().max()) for v in g64.values())
        pre = "m%d_" % len(meta)
        store[pre + "pred"] = p64.numpy()
        store[pre + "loss"] = np.array(l64, dtype=np.float64)
        own = {"pred": rel(p32, p64, 0.0), "loss": abs(l32 - l64) / max(abs(l64), 1e-6)}
        for k in ("x", "edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr",
                  "batch", "y"):
            v = b._store[k].contiguous().numpy()
            store[pre + "b_" + k] = v.astype(np.int32) if v.dtype == np.int64 else v
        for k, v in sd.items():
            store[pre + "sd_" + k] = v.numpy()
        for k, g in g64.items():
            store[pre + "gp_" + k] = g.numpy().astype(np.float32) if g.numel() > 4096 else g.numpy()
            own["gp_" + k] = rel(g32[k], g, 1e-2 * gmax)
        meta.append({"name": name, "num_graphs": len(graphs), "own32": own})
        print("  models_cfg: %s loss %.6f, worst own fp32 error %.2e" % (name, l64, max(own.values())))
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "models_cfg.npz"), **store)


def main():
    if not refimport.available():
        raise SystemExit("reference tree not found at %s" % refimport.REF_ROOT)
    os.makedirs(OUT, exist_ok=True)
    ns = refimport.load()
    only = sys.argv[1:]
    for name, fn in (("extract", make_extract), ("extract_full", make_extract_full), ("layers", make_layers),
                     ("model", make_model), ("models_cfg", make_models_cfg)):
        if not only or name in only:
            fn(ns)


if __name__ == "__main__":
    main()
