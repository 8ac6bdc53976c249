"""Fused graph-regression head (csrc/head.cu) against the same stage in plain torch evaluated in float64: ReLU, per-graph
sum / mean pooling, Linear(H,1), L1 / MSE loss (models/GNNs.py:276-277, GraphRegression.py:26, train_ZINC.py:42) --
loss, scores and every gradient; padded capacity batches; and the whole model: fused_loss == loss_fn(model(batch))."""
import pytest
import torch

from tests.util import RTOL, rel_err

pytestmark = pytest.mark.gpu


def _case(N, H, G, seed, empty_graph=False):
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(1, max(2, 2 * N // G), (G,), generator=g)
    if empty_graph:
        sizes[G // 2] = 0
    batch = torch.repeat_interleave(torch.arange(G), sizes)
    N = int(batch.numel())
    rep = torch.randn(N, H, generator=g)
    rep[rep.abs() < 1e-3] = 0.5                       # keep the draw clear of the ReLU kink
    w = torch.randn(1, H, generator=g) * 0.2
    b = torch.randn(1, generator=g)
    y = torch.randn(G, generator=g)
    return rep, w, b, y, batch


@pytest.mark.parametrize("N,H,G", [(300, 104, 16), (2986, 104, 128), (50, 20, 7), (4000, 256, 33)])
@pytest.mark.parametrize("mean", [False, True])
@pytest.mark.parametrize("kind", ["l1", "mse"])
def test_fused_head_matches_torch(lib, N, H, G, mean, kind):
    from kpgnn_b200.head import fused_regression_loss
    dev = torch.device("cuda:0")
    rep0, w0, b0, y, batch = _case(N, H, G, N + H, empty_graph=(G == 33))
    outs = []
    for mode in ("ref", "fused"):
        dt = torch.float64 if mode == "ref" else torch.float32
        rep = rep0.to(dev, dt).requires_grad_(True)
        w, b = w0.to(dev, dt).requires_grad_(True), b0.to(dev, dt).requires_grad_(True)
        yy, bt = y.to(dev, dt), batch.to(dev)
        if mode == "ref":
            h = torch.relu(rep)
            pooled = torch.zeros(G, H, dtype=dt, device=dev).index_add_(0, bt, h)
            if mean:
                cnt = torch.bincount(bt, minlength=G).clamp(min=1).to(dt).unsqueeze(1)
                pooled = pooled / cnt
            score = (pooled @ w.t()).squeeze(1) + b
            loss = (score - yy).abs().mean() if kind == "l1" else ((score - yy) ** 2).mean()
        else:
            loss, score = fused_regression_loss(rep, w, b, yy, bt, G, mean=mean, kind=kind, return_score=True)
        (loss * 1.7).backward()
        outs.append((loss.detach(), score.detach(), rep.grad, w.grad, b.grad))
    names = ("loss", "score", "drep", "dw", "db")
    for n, a, c in zip(names, outs[0], outs[1]):
        assert rel_err(c, a) < RTOL, (n, rel_err(c, a))
    # bit-reproducible
    rep = rep0.to(dev).requires_grad_(True)
    l2, s2 = fused_regression_loss(rep, w0.to(dev), b0.to(dev), y.to(dev), batch.to(dev), G, mean=mean, kind=kind,
                                   return_score=True)
    assert torch.equal(l2, outs[1][0]) and torch.equal(s2, outs[1][1])


def test_fused_head_padded_rows(lib):
    """Rows behind n_dev (padding of a capacity batch, graph id G) contribute nothing and get zero gradient."""
    from kpgnn_b200.head import fused_regression_loss
    dev = torch.device("cuda:0")
    rep0, w0, b0, y, batch = _case(500, 104, 20, 3)
    N, cap = rep0.size(0), rep0.size(0) + 37
    res = []
    for padded in (False, True):
        rep = rep0.to(dev)
        bt = batch.to(dev)
        n_dev = None
        if padded:
            rep = torch.cat([rep, torch.randn(cap - N, 104, device=dev)])          # garbage in the padding rows
            bt = torch.cat([bt, torch.full((cap - N,), 20, dtype=torch.int64, device=dev)])
            n_dev = torch.tensor([N], dtype=torch.int32, device=dev)
        rep = rep.requires_grad_(True)
        w, b = w0.to(dev).requires_grad_(True), b0.to(dev).requires_grad_(True)
        loss = fused_regression_loss(rep, w, b, y.to(dev), bt, 20, n_dev=n_dev)
        loss.backward()
        res.append((loss.detach(), rep.grad, w.grad, b.grad))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2]) and torch.equal(res[0][3], res[1][3])
    assert torch.equal(res[0][1], res[1][1][:N]) and float(res[1][1][N:].abs().max()) == 0.0


def test_model_fused_loss_equals_unfused(lib):
    from kpgnn_b200.model import l1_loss, zinc_kpginplus
    from tests.test_varbatch_gpu import _host_batch
    dev = torch.device("cuda:0")
    hb = _host_batch(12, 41)
    batch = hb.to(dev) if hasattr(hb, "to") else hb
    torch.manual_seed(0)
    model = zinc_kpginplus(8, 8, 104).to(dev).train()
    res = []
    for fused in (False, True):
        model.zero_grad(set_to_none=True)
        loss = model.fused_loss(batch, "l1") if fused else l1_loss(model(batch), batch.y)
        assert loss is not None
        loss.backward()
        res.append((loss.detach(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    assert abs(float(res[0][0]) - float(res[1][0])) <= 1e-6 * abs(float(res[0][0]))
    gmax = max(float(g.abs().max()) for g in res[0][1].values())
    for n, g in res[0][1].items():
        assert rel_err(res[1][1][n], g, floor=1e-3 * gmax) < 2e-5, (n, rel_err(res[1][1][n], g, floor=1e-3 * gmax))
