"""CPU: pins the oracle (oracle/extract_np.py, oracle/layers_torch.py, oracle/model_torch.py) against the
reference's own outputs committed under tests/golden/, and -- when /root/reference is present (build container
only) -- against the reference executed live."""
import numpy as np
import pytest
import torch

from oracle import layers_torch as OL
from oracle import refimport
from oracle.extract_np import extract_multi_hop_neighbors_np
from oracle.model_torch import l1_loss, zinc_oracle_model
from tests import golden_util as GU
from tests.util import rel_err

ORACLE_LAYERS = {"KPGINConv": OL.OracleKPGINConv, "KPGINPlusConv": OL.OracleKPGINPlusConv,
                 "KPGCNConv": OL.OracleKPGCNConv, "KPGraphSAGEConv": OL.OracleKPGraphSAGEConv,
                 "GINEConv": OL.OracleGINEConv}


def test_extraction_oracle_matches_reference_goldens():
    z, meta = GU.load("extract.npz")
    assert len(meta) >= 20
    for i, m in enumerate(meta):
        pre = "c%d_" % i
        ea = z[pre + "in_edge_attr"] if m["typed"] else None
        out = extract_multi_hop_neighbors_np(m["num_nodes"], z[pre + "in_edge_index"], ea, *m["args"])
        got_fields = sorted(k for k, v in out.items() if v is not None)
        assert got_fields == sorted(m["fields"]), (m["name"], got_fields, m["fields"])
        for f in m["fields"]:
            ref = z[pre + "out_" + f]
            assert out[f].shape == ref.shape, (m["name"], f)
            assert np.array_equal(out[f], ref), (m["name"], f)      # integer work: bit-exact


def test_layer_oracle_matches_reference_goldens():
    z, meta = GU.load("layers.npz")
    for i, m in enumerate(meta):
        c = GU.layer_case(z, i)
        layer = GU.build_layer(m["ctor"], ORACLE_LAYERS)
        layer.load_state_dict(c["sd"])
        layer.train()
        x = c["x"].clone().requires_grad_(True)
        P = c["P"].clone().requires_grad_(True) if "P" in c else None
        if m["gine"]:
            y = layer(x * 1.0, c["edge_index"], c["edge_attr"][:, :1])
        else:
            y = layer(x * 1.0, c["edge_index"], c["edge_attr"], c.get("pe"), P)
        y.backward(c["gy"])
        assert rel_err(y, c["y"]) < 1e-6, m["name"]
        assert rel_err(x.grad, c["gx"]) < 1e-5, m["name"]
        if P is not None:
            assert rel_err(P.grad, c["gP"]) < 1e-5, m["name"]
        gmax = max(float(v.abs().max()) for v in c["gp"].values())
        for n, p in layer.named_parameters():
            if n in c["gp"]:
                # a bias feeding a BatchNorm has an analytically zero gradient: what is stored is BLAS
                # rounding noise (~1e-7 of the layer's gradient scale) that differs from host to host, so
                # such tensors are held to 1e-6 of the layer's largest gradient instead of to themselves
                assert rel_err(p.grad, c["gp"][n], floor=1e-1 * gmax) < 1e-5, (m["name"], n)


def test_model_oracle_matches_reference_golden():
    z, _ = GU.load("model_zinc.npz")
    model = zinc_oracle_model()
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_")})
    model.train()
    b = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("b_")}
    b["num_graphs"] = int(b["batch"].max()) + 1
    score = model(b)
    loss = l1_loss(score, b["y"])
    loss.backward()
    assert rel_err(score, torch.from_numpy(z["score"])) < 1e-6
    assert abs(loss.item() - float(z["loss"])) < 1e-6 * abs(float(z["loss"]))
    gmax = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith("gp_"))
    for n, p in model.named_parameters():
        assert rel_err(p.grad, torch.from_numpy(z["gp_" + n]), floor=1e-2 * gmax) < 1e-5, n


@pytest.mark.skipif(not refimport.available(), reason="reference tree only exists in the build container")
def test_extraction_oracle_matches_live_reference():
    from kpgnn_b200 import synth
    ns = refimport.load()
    rng = np.random.default_rng(77)
    cases = []
    for g in synth.zinc_like_graphs(3, seed=77):
        cases.append((g, (8, 50, 6, 3, 50, 50, "spd")))
        cases.append((g, (3, 2, 3, 2, 4, 4, "gd")))
    for i in range(6):
        g = synth.random_typed_graph(rng, int(rng.integers(3, 20)), 0.3, directed=bool(i % 2), typed=bool(i % 3))
        if g["edge_index"].shape[1]:
            cases.append((g, (int(rng.integers(1, 5)), 3, int(rng.integers(0, 3)), int(rng.integers(0, 3)), 3, 3,
                              "spd" if i % 2 else "gd")))
    for g, args in cases:
        d = ns.Data(x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"]),
                    edge_attr=None if g["edge_attr"] is None else torch.from_numpy(g["edge_attr"]))
        d.num_nodes_ = g["num_nodes"]
        r = ns.data_utils.extract_multi_hop_neighbors(d, *args)
        o = extract_multi_hop_neighbors_np(g["num_nodes"], g["edge_index"], g["edge_attr"], *args)
        for k, v in o.items():
            rv = r._store.get(k, None)
            if v is None:
                assert rv is None
            else:
                assert np.array_equal(rv.numpy(), v), (k, args)


@pytest.mark.skipif(not refimport.available(), reason="reference tree only exists in the build container")
def test_dropin_layers_construct_inside_unmodified_reference_backbone():
    """The drop-in classes satisfy everything models/GNNs.py asks of a layer at construction time (attributes,
    deepcopy, ModuleList, reset_parameters) and yield the reference's exact state_dict key set.  Forward needs a
    GPU and is covered by tests/test_golden_gpu.py."""
    import argparse
    import sys
    ns = refimport.load()
    from kpgnn_b200.layers import layer_utils as mine
    common = dict(JK="concat", norm_type="Batch", residual=True, virtual_node=False, use_rd=False, num_hop1_edge=3,
                  max_edge_count=50, max_hop_num=6, max_distance_count=50, wo_peripheral_edge=False,
                  wo_peripheral_configuration=False, drop_prob=0.0)
    for name, cls in (("KPGINPlus", ns.GNNs.GNNPlus), ("KPGIN", ns.GNNs.GNN), ("KPGINPrime", ns.GNNs.GNNPrime),
                      ("KPGCN", ns.GNNs.GNN), ("KPGraphSAGE", ns.GNNs.GNN)):
        args = argparse.Namespace(model_name=name, hidden_size=96, K=8, num_hop1_edge=3, max_pe_num=50,
                                  combine="geometric", num_layer=8, eps=0., train_eps=False, aggr="add")
        keys = []
        for factory in (ns.layer_utils.make_gnn_layer, mine.make_gnn_layer):
            gnn = cls(num_layer=8, gnn_layer=factory(args), init_emb=ns.input_encoder.EmbeddingEncoder(21, 96),
                      **common)
            keys.append(list(gnn.state_dict().keys()))
            shapes = {k: tuple(v.shape) for k, v in gnn.state_dict().items()}
        assert keys[0] == keys[1], name
    assert "kpgnn_b200.layers.KPGINplus" in sys.modules


def test_install_dropin_aliases_reference_import_names():
    import subprocess
    import sys
    code = ("import kpgnn_b200, sys; kpgnn_b200.install_dropin();"
            "from layers.gine import GINEConv; from layers.feature_encoder import FeatureConcatEncoder;"
            "from layers.layer_utils import make_gnn_layer; from layers.combine import *;"
            "from data_utils import extract_multi_hop_neighbors, resistance_distance, post_transform;"
            "assert GINEConv.__module__ == 'kpgnn_b200.layers.gine'; assert nn is not None and torch is not None;"
            "print('ok')")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


@pytest.mark.skipif(not refimport.available(), reason="reference files not staged")
def test_reference_models_import_over_dropin_namespace():
    """oracle/refimport.load_models_over_dropin: the reference's models/GNNs.py imported unmodified with its
    `from layers...` lines bound to the product package, next to the all-reference namespace (CPU: construction only;
    forward/backward on a B200 is tests/test_reference_models_gpu.py)."""
    from tests import ref_util as RU
    from kpgnn_b200.layers import layer_utils as mine
    from kpgnn_b200.layers.input_encoder import EmbeddingEncoder
    dropin = refimport.load_models_over_dropin()
    for name, cfg in RU.CONFIGS.items():
        m = RU.build_model(cfg, dropin.GNNs, mine.make_gnn_layer, EmbeddingEncoder, dropin)
        r = RU.build_model(cfg, dropin.ref.GNNs, dropin.ref.layer_utils.make_gnn_layer,
                           dropin.ref.input_encoder.EmbeddingEncoder, dropin.ref)
        assert list(m.state_dict().keys()) == list(r.state_dict().keys()), name
        assert type(RU.first_layer(m)).__module__.startswith("kpgnn_b200.layers.")
        assert not type(RU.first_layer(r)).__module__.startswith("kpgnn_b200")
