"""BENCH INFRASTRUCTURE ONLY (the `--impl reference` arm and the cpu_baseline leg of bench.py): the reference's OWN
training step for BASELINE.json configs[1] -- `train()` of train_ZINC.py:29-47 over `get_model()` of train_ZINC.py:50-83
-- built from the UNMODIFIED reference files (oracle/refimport.py: /root/reference here, the staged copy oracle/_ref/ on
the GPU box) behind the torch_geometric stand-in, on the host's CPU cores.  No product code is involved: extraction is
the reference's `extract_multi_hop_neighbors` per graph, collation is `Batch.from_data_list`."""
import argparse
import os

import numpy as np
import torch


def available():
    from oracle import refimport
    return refimport.available()


def zinc_reference_trainer(graphs, K=8, num_layer=8, hidden=104, extract_args=None, lr=1e-3, threads=None):
    """Returns (step, info): step() runs one optimisation step on one pre-collated batch and returns the loss."""
    from oracle import refimport
    ns = refimport.load()
    torch.set_num_threads(threads or os.cpu_count())
    extract_args = extract_args or (K, 50, 6, 3, 50, 50, "spd")            # train_ZINC.py:124-134
    datas = []
    for g in graphs:
        d = ns.Data(x=torch.from_numpy(np.asarray(g["x"])), edge_index=torch.from_numpy(np.asarray(g["edge_index"])),
                    edge_attr=torch.from_numpy(np.asarray(g["edge_attr"])))
        d.num_nodes_ = g["num_nodes"]
        d = ns.data_utils.extract_multi_hop_neighbors(d, *extract_args)
        d.y = torch.tensor([g["y"]], dtype=torch.float32)
        datas.append(d)
    batch = ns.Batch.from_data_list(datas)
    torch.manual_seed(0)
    args = argparse.Namespace(model_name="KPGINPlus", hidden_size=hidden, K=K, num_hop1_edge=3, max_pe_num=50,
                              combine="geometric", num_layer=num_layer, eps=0., train_eps=False, aggr="add")
    gnn = ns.GNNs.GNNPlus(num_layer=num_layer, gnn_layer=ns.layer_utils.make_gnn_layer(args), JK="concat",
                          norm_type="Batch", init_emb=ns.input_encoder.EmbeddingEncoder(21, hidden), residual=True,
                          virtual_node=False, use_rd=False, num_hop1_edge=3, max_edge_count=50, max_hop_num=6,
                          max_distance_count=50, wo_peripheral_edge=False, wo_peripheral_configuration=False,
                          drop_prob=0.0)
    model = ns.GraphRegression.GraphRegression(embedding_model=gnn, pooling_method="sum")
    model.reset_parameters()
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=lr)                       # train_ZINC.py:244

    def step():
        opt.zero_grad()
        score = model(batch).squeeze()
        loss = (score - batch.y).abs().mean()                               # train_ZINC.py:42
        loss.backward()
        opt.step()
        return float(loss.item())

    info = {"kind": "reference", "threads": torch.get_num_threads(), "nodes": int(batch.num_nodes),
            "khop_edges": int(batch.edge_index.size(1)),
            "params": sum(p.numel() for p in model.parameters())}
    return step, info
