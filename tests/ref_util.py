"""Helpers for the tests that run the UNMODIFIED reference (oracle/refimport.py: /root/reference in the build container,
the staged copy oracle/_ref/ on the GPU box) next to the product."""
import argparse
import os
import pickle

import numpy as np
import torch

from oracle import refimport

# BASELINE.json configs -> (model_name, GNN class, extraction args, model hyper-parameters); train_*.py of the reference
CONFIGS = {
    # configs[0]: train_EXP.py:148-158 + its argparse defaults (hidden 48, K 3, 3 layers, JK last, sum pooling)
    "exp": dict(model_name="KPGIN", cls="GNN", extract=(3, 1, 5, 1, 1000, 1000, "spd"), hidden_size=48, K=3,
                num_layer=3, num_hop1_edge=1, max_pe_num=1, max_edge_count=1000, max_hop_num=5,
                max_distance_count=1000, JK="last", residual=False, input_size=2, head=("classification", 2)),
    # configs[1]: train_ZINC.py:124-134, README.md:127
    "zinc": dict(model_name="KPGINPlus", cls="GNNPlus", extract=(8, 50, 6, 3, 50, 50, "spd"), hidden_size=104, K=8,
                 num_layer=8, num_hop1_edge=3, max_pe_num=50, max_edge_count=50, max_hop_num=6,
                 max_distance_count=50, JK="concat", residual=True, input_size=21, head=("regression", 1)),
    # configs[2]: README.md:128 (K=16, hidden 96; K-hop only at layer 1, then GINE layers); 5 layers here
    "prime": dict(model_name="KPGINPrime", cls="GNNPrime", extract=(16, 50, 6, 3, 50, 50, "spd"), hidden_size=96, K=16,
                  num_layer=5, num_hop1_edge=3, max_pe_num=50, max_edge_count=50, max_hop_num=6,
                  max_distance_count=50, JK="concat", residual=True, input_size=21, head=("regression", 1)),
    # configs[3]: train_SR.py:115-125 + its argparse defaults (hidden 48, K 4, max_pe_num 1000), gd kernel
    "sr_gcn": dict(model_name="KPGCN", cls="GNN", extract=(4, 1000, 4, 1, 1000, 1000, "gd"), hidden_size=48, K=4,
                   num_layer=4, num_hop1_edge=1, max_pe_num=1000, max_edge_count=1000, max_hop_num=4,
                   max_distance_count=1000, JK="last", residual=False, input_size=2, head=("classification", 15)),
    "sr_sage": dict(model_name="KPGraphSAGE", cls="GNN", extract=(4, 1000, 4, 1, 1000, 1000, "gd"), hidden_size=48,
                    K=4, num_layer=4, num_hop1_edge=1, max_pe_num=1000, max_edge_count=1000, max_hop_num=4,
                    max_distance_count=1000, JK="last", residual=False, input_size=2, head=("classification", 15)),
}


def available():
    return refimport.available()


def exp_graphs(n=128):
    """First n graphs of the reference's in-repo EXP dataset as raw-graph dicts (train_EXP.py:63-65: x = x[:,0].long())."""
    refimport.load()          # registers the stand-in Data class the pickle refers to
    with open(os.path.join(refimport.REF_ROOT, "data", "EXP", "raw", "GRAPHSAT.pkl"), "rb") as f:
        exp = pickle.load(f)
    out = []
    for d in exp[:n]:
        out.append({"num_nodes": int(d.x.size(0)), "x": d.x[:, 0].long().numpy(), "edge_index": d.edge_index.numpy(),
                    "edge_attr": None, "y": int(d.y.view(-1)[0])})
    return out


def sr25_graphs(random_x=False):
    """The 15 in-repo strongly regular graphs srg(25,12,5,6) (datasets/SRDataset.py): x = 1 as long.
    random_x: seeded node types in {0, 1} instead -- with constant features every node of a strongly regular graph
    carries the same embedding, BatchNorm then divides by a near-zero batch variance and turns fp32 rounding into
    1e-3 relative noise on both sides; parity fixtures need a well-conditioned input."""
    import networkx as nx
    gs = nx.read_graph6(os.path.join(refimport.REF_ROOT, "data", "sr25", "raw", "sr251256.g6"))
    out = []
    for i, g in enumerate(gs):
        e = np.array(list(g.to_directed().edges)).T
        e = e[:, np.lexsort((e[1], e[0]))]
        x = np.random.default_rng(100 + i).integers(0, 2, size=25) if random_x else np.ones(25, dtype=np.int64)
        out.append({"num_nodes": 25, "x": x.astype(np.int64), "edge_index": e.astype(np.int64),
                    "edge_attr": None, "y": i})
    return out


def ref_batch(ns, graphs, extract_args, y_dtype=torch.float32):
    """Reference extraction per graph (data_utils.py, unmodified) + Batch.from_data_list collation."""
    datas = []
    for g in graphs:
        d = ns.Data(x=torch.from_numpy(np.asarray(g["x"])), edge_index=torch.from_numpy(np.asarray(g["edge_index"])),
                    edge_attr=None if g["edge_attr"] is None else torch.from_numpy(np.asarray(g["edge_attr"])))
        d.num_nodes_ = g["num_nodes"]
        datas.append(ns.data_utils.extract_multi_hop_neighbors(d, *extract_args))
    b = ns.Batch.from_data_list(datas)
    b.y = torch.tensor([g.get("y", 0.0) for g in graphs], dtype=y_dtype)
    return b


def build_model(cfg, gnns_mod, make_gnn_layer, init_emb_cls, head_mods, combine="geometric", virtual_node=False,
                drop_prob=0.0, norm_type="Batch"):
    """get_model() of the reference's train scripts (train_ZINC.py:50-83, train_EXP.py / train_SR.py get_model)."""
    args = argparse.Namespace(model_name=cfg["model_name"], hidden_size=cfg["hidden_size"], K=cfg["K"],
                              num_hop1_edge=cfg["num_hop1_edge"], max_pe_num=cfg["max_pe_num"], combine=combine,
                              num_layer=cfg["num_layer"], eps=0., train_eps=False, aggr="add")
    layer = make_gnn_layer(args)
    gnn = getattr(gnns_mod, cfg["cls"])(
        num_layer=cfg["num_layer"], gnn_layer=layer, JK=cfg["JK"], norm_type=norm_type,
        init_emb=init_emb_cls(cfg["input_size"], cfg["hidden_size"]), residual=cfg["residual"],
        virtual_node=virtual_node, use_rd=False, num_hop1_edge=cfg["num_hop1_edge"],
        max_edge_count=cfg["max_edge_count"], max_hop_num=cfg["max_hop_num"],
        max_distance_count=cfg["max_distance_count"], wo_peripheral_edge=False, wo_peripheral_configuration=False,
        drop_prob=drop_prob)
    kind, out = cfg["head"]
    if kind == "regression":
        model = head_mods.GraphRegression.GraphRegression(embedding_model=gnn, pooling_method="sum")
    else:
        model = head_mods.GraphClassification.GraphClassification(embedding_model=gnn, pooling_method="sum",
                                                                 output_size=out)
    model.reset_parameters()
    return model


def loss_fn(cfg, pred, y):
    if cfg["head"][0] == "regression":
        return (pred.squeeze() - y.squeeze()).abs().mean()                              # train_ZINC.py:42
    return torch.nn.functional.nll_loss(torch.log_softmax(pred, dim=-1), y.long())      # train_EXP.py:40-42


def first_layer(model):
    em = model.embedding_model
    return (em.gnns if hasattr(em, "gnns") else em.khop_gnns)[0]
