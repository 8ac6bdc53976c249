"""GPU parity of the product path against the REFERENCE's own outputs (tests/golden/*.npz, produced by running
/root/reference unmodified on the CPU in the build container, oracle/make_golden.py).
Bar: 1e-5 relative, fp32 (north_star); gradients get 5e-5 because GPU GEMM/BatchNorm reductions downstream of
the kernels run in a different order than the CPU reference's."""
import numpy as np
import pytest
import torch

from tests import golden_util as GU
from tests.util import rel_err

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
NOISE_ONLY = ("mlp.0.bias", "mlp.3.bias")


def _product_layers():
    from kpgnn_b200.layers.KPGIN import KPGINConv
    from kpgnn_b200.layers.KPGINplus import KPGINPlusConv
    from kpgnn_b200.layers.KPGCN import KPGCNConv
    from kpgnn_b200.layers.KPGraphSAGE import KPGraphSAGEConv
    from kpgnn_b200.layers.gine import GINEConv
    return {"KPGINConv": KPGINConv, "KPGINPlusConv": KPGINPlusConv, "KPGCNConv": KPGCNConv,
            "KPGraphSAGEConv": KPGraphSAGEConv, "GINEConv": GINEConv}


_Z, _META = GU.load("layers.npz")


@pytest.mark.parametrize("idx", range(len(_META)), ids=[m["name"] for m in _META])
def test_layer_matches_reference_golden(lib, idx):
    m = _META[idx]
    c = GU.layer_case(_Z, idx)
    dev = torch.device("cuda:0")
    layer = GU.build_layer(m["ctor"], _product_layers())
    layer.load_state_dict(c["sd"])              # the reference's state_dict loads unchanged
    layer = layer.to(dev).train()
    x = c["x"].to(dev).requires_grad_(True)
    P = c["P"].to(dev).requires_grad_(True) if "P" in c else None
    ei, ea = c["edge_index"].to(dev), c["edge_attr"].to(dev)
    if m["gine"]:
        y = layer(x * 1.0, ei, ea[:, :1])
    else:
        pe = c["pe"].to(dev) if "pe" in c else None
        y = layer(x * 1.0, ei, ea, pe, P)
    y.backward(c["gy"].to(dev))
    assert rel_err(y, c["y"]) < 1e-5, ("forward", rel_err(y, c["y"]))
    assert rel_err(x.grad, c["gx"]) < 5e-5, ("dx", rel_err(x.grad, c["gx"]))
    if P is not None:
        assert rel_err(P.grad, c["gP"]) < 5e-5, ("dP", rel_err(P.grad, c["gP"]))
    gmax = max(float(v.abs().max()) for v in c["gp"].values())
    for n, p in layer.named_parameters():
        if n not in c["gp"]:
            continue
        ref = c["gp"][n]
        if p.grad is None:
            assert float(ref.abs().max()) == 0.0, n
        elif n in NOISE_ONLY:
            assert float((p.grad.cpu() - ref).abs().max()) < 1e-4 * gmax, n
        else:
            err = rel_err(p.grad, ref, floor=1e-3 * gmax)
            assert err < 5e-5, (n, err)


def test_model_matches_reference_golden(lib):
    """GraphRegression(GNNPlus(KPGINPlus K=8 L=8 H=104)): the reference's state_dict loads into the product
    backbone unchanged; score, loss and every parameter gradient match the reference's."""
    from kpgnn_b200.model import Batch, l1_loss, zinc_kpginplus
    z, _ = GU.load("model_zinc.npz")
    dev = torch.device("cuda:0")
    model = zinc_kpginplus()
    missing = model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_")})
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(dev).train()
    fields = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("b_")}
    b = Batch(num_graphs=int(fields["batch"].max()) + 1, **fields).to(dev)
    score = model(b)
    loss = l1_loss(score, b.y)
    loss.backward()
    assert rel_err(score, torch.from_numpy(z["score"])) < 1e-5
    assert abs(loss.item() - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    gmax = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith("gp_"))
    worst = 0.0
    for n, p in model.named_parameters():
        ref = torch.from_numpy(z["gp_" + n])
        g = p.grad if p.grad is not None else torch.zeros_like(ref)
        worst = max(worst, rel_err(g, ref, floor=1e-2 * gmax))
    assert worst < 5e-5, worst


def test_determinism_bitwise(lib):
    """Two runs of forward+backward on the same inputs give bit-identical outputs and gradients (no float
    atomics anywhere on the K-hop path)."""
    from kpgnn_b200.layers.KPGINplus import KPGINPlusConv
    from tests.util import zinc_batch
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    b = zinc_batch(16, 8, "spd", seed=9)
    N = b["num_nodes"]
    layer = KPGINPlusConv(104, 104, 8, 3, 50, "geometric").to(dev).train()
    x0 = torch.randn(N, 8, 104, device=dev)
    P0 = torch.randn(N, 8, 104, device=dev)
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    from kpgnn_b200.ops import khop_aggregate, ACT_GELU
    from kpgnn_b200.plan import get_plan
    res = []
    for _ in range(2):
        x = x0.clone().requires_grad_(True)
        P = P0.clone().requires_grad_(True)
        t0 = layer.hop1_edge_emb.weight.detach().clone().requires_grad_(True)
        tk = layer.hopk_edge_emb.weight.detach().clone().requires_grad_(True)
        th = torch.softmax(torch.randn(8, 104, device=dev, generator=torch.Generator(dev).manual_seed(1)), 0)
        th.requires_grad_(True)
        ei2 = ei.clone()   # fresh plan each time
        plan, k = get_plan(ei2, ea, N)
        y = khop_aggregate(x, plan, k, P=P, T0=t0, Tk=tk, theta=th, act=ACT_GELU, fuse=True)
        y.backward(torch.ones_like(y))
        res.append([t.detach().clone() for t in (y, x.grad, P.grad, t0.grad, tk.grad, th.grad)])
    for a, c in zip(*res):
        assert torch.equal(a, c)


@pytest.mark.parametrize("gate,H", [("tanh", 104), ("sigmoid", 16), ("sigmoid", 12), ("tanh", 96)])
def test_fused_peripheral_encoder_matches_reference_form(lib, gate, H):
    """kp_table_sum_* (folded tables + gather-sum) against the reference's lookup -> concat -> Linear -> sum
    (models/GNNs.py:393-400), forward and every parameter gradient."""
    from kpgnn_b200.encoders import fused_peripheral_attr, peripheral_index
    from kpgnn_b200.layers.feature_encoder import FeatureConcatEncoder
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    N, K, c = 333, 8, 3
    sq = torch.tanh if gate == "tanh" else torch.sigmoid
    ee = FeatureConcatEncoder([5, 51], H, padding=0).to(dev)
    ce = FeatureConcatEncoder([51] * 7, H, padding=0).to(dev)
    pew = torch.randn(1, device=dev, requires_grad=True)
    pcw = torch.randn(1, device=dev, requires_grad=True)
    pea = torch.stack([torch.randint(0, 5, (N, K, c)), torch.randint(0, 51, (N, K, c))], -1).to(dev)
    pca = torch.randint(0, 51, (N, K, 7)).to(dev)
    gy = torch.randn(N, K, H, device=dev)
    res = []
    for fused in (False, True):
        for p in list(ee.parameters()) + list(ce.parameters()) + [pew, pcw]:
            p.grad = None
        if fused:
            P = fused_peripheral_attr(ee, ce, pew, pcw, peripheral_index(pea, pca), N, K, c, gate=gate)
        else:
            P = sq(pew) * ee(pea).sum(-2) + sq(pcw) * ce(pca)
        P.backward(gy)
        res.append([P.detach()] + [p.grad.clone() for p in list(ee.parameters()) + list(ce.parameters()) + [pew, pcw]])
    for a, b in zip(res[1], res[0]):
        assert rel_err(a, b) < 2e-5, rel_err(a, b)


_ZC, _METAC = GU.load("models_cfg.npz")


@pytest.mark.parametrize("idx", range(len(_METAC)), ids=[m["name"] for m in _METAC])
def test_other_config_models_match_reference_golden(lib, idx):
    """BASELINE.json configs[0], [2], [3] (EXP KP-GIN K=3 H=48, KPGINPrime K=16 H=96, SR25 KP-GCN / KP-GraphSAGE gd K=4
    with 1 002-row tables): the reference's state_dict loads into the product backbones (kpgnn_b200/backbones.py)
    unchanged; prediction, loss and every parameter gradient match what the reference computed on the CPU."""
    from kpgnn_b200 import backbones
    from kpgnn_b200.model import Batch
    from tests import ref_util as RU
    m = _METAC[idx]
    cfg = RU.CONFIGS[m["name"]]
    pre = "m%d_" % idx
    dev = torch.device("cuda:0")
    model = backbones.make_model(cfg["model_name"], cfg["hidden_size"], cfg["K"], cfg["num_layer"], cfg["input_size"],
                                 cfg["num_hop1_edge"], cfg["max_pe_num"], cfg["max_edge_count"], cfg["max_hop_num"],
                                 cfg["max_distance_count"], JK=cfg["JK"], residual=cfg["residual"],
                                 output_size=None if cfg["head"][0] == "regression" else cfg["head"][1])
    res = model.load_state_dict({k[len(pre) + 3:]: torch.from_numpy(_ZC[k]) for k in _ZC.files
                                 if k.startswith(pre + "sd_")})
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(dev).train()
    fields = {}
    for k in _ZC.files:
        if k.startswith(pre + "b_"):
            v = torch.from_numpy(_ZC[k])
            fields[k[len(pre) + 2:]] = v.long() if v.dtype == torch.int32 else v
    b = Batch(num_graphs=m["num_graphs"], **fields).to(dev)
    pred = model(b)
    loss = RU.loss_fn(cfg, pred, b.y)
    loss.backward()
    own = m["own32"]            # the reference's own fp32-vs-float64 distance per tensor (oracle/make_golden.py)

    def bar(key):
        return max(1e-5, 10.0 * own.get(key, 0.0))
    assert rel_err(pred, torch.from_numpy(_ZC[pre + "pred"])) < bar("pred")
    assert abs(loss.item() - float(_ZC[pre + "loss"])) < bar("loss") * abs(float(_ZC[pre + "loss"]))
    gmax = max(float(np.abs(_ZC[k]).max()) for k in _ZC.files if k.startswith(pre + "gp_"))
    for n, p in model.named_parameters():
        key = pre + "gp_" + n
        if key not in _ZC.files:
            assert p.grad is None or float(p.grad.abs().max()) <= 1e-6 * gmax, n
            continue
        ref = torch.from_numpy(_ZC[key])
        g = p.grad if p.grad is not None else torch.zeros_like(ref)
        if n.endswith(NOISE_ONLY + ("combine_proj.bias",)):      # a bias feeding BatchNorm: analytically zero gradient
            assert float((g.cpu().double() - ref.double()).abs().max()) < 1e-4 * gmax, n
            continue
        err = rel_err(g, ref, floor=1e-2 * gmax)
        assert err < bar("gp_" + n), (n, err, own.get("gp_" + n))
