"""Batches of different sizes through static buffers: the compact wire format and its device-side
unpacking, the dense block's device-side row count, and a captured training step replayed over distinct batches --
loss trajectory against the oracle model stepping on the same (unpadded) batches with torch.optim.Adam."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from kpgnn_b200 import synth
from tests.util import collate, rel_err

pytestmark = pytest.mark.gpu
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

ARGS = (8, 50, 6, 3, 50, 50, "spd")


def _host_batch(num_graphs, seed):
    from kpgnn_b200.model import Batch
    graphs = synth.zinc_like_graphs(num_graphs, seed=seed)
    d = collate(graphs, ARGS)                     # oracle extraction (numpy), reference collation
    return Batch(**d)


def test_wire_pack_unpack_roundtrip(lib):
    from kpgnn_b200.train import fit_spec
    from kpgnn_b200.wire import DeviceWire
    dev = torch.device("cuda:0")
    hbs = [_host_batch(12, s) for s in (1, 2, 3)]
    spec, bounds = fit_spec(hbs, 8, 3, 6)
    assert spec.attr_bytes == 1 and spec.p_bytes == 1 and spec.nbytes < sum(hbs[0].nbytes() for _ in range(1))
    w = DeviceWire(spec, dev)
    flat = spec.host_buffer()
    for hb in hbs:
        spec.pack(hb, flat)
        w.stage.copy_(flat)
        b = w.unpack()
        N, E = hb.x.size(0), hb.edge_index.size(1)
        assert int(w.n_dev) == N
        for f in ("x", "edge_attr", "peripheral_edge_attr", "peripheral_configuration_attr", "batch"):
            got, ref = getattr(b, f).cpu(), getattr(hb, f)
            n = E if f == "edge_attr" else N
            assert torch.equal(got[:n], ref), f
            pad = got[n:]
            assert bool((pad == (spec.G if f == "batch" else 0)).all()), f
        assert torch.equal(b.edge_index.cpu()[:, :E], hb.edge_index) and int(b.edge_index[:, E:].abs().sum()) == 0
        assert torch.equal(b.y.cpu(), hb.y)
    with pytest.raises(ValueError):
        spec.pack(_host_batch(13, 4), flat)       # wrong graph count


@pytest.mark.parametrize("N,cap", [(2952, 3200), (300, 301), (37, 5000), (1000, 1000)])
def test_dense_block_device_row_count(lib, N, cap):
    """kp_dense_desc.n_dev: a [cap, C] buffer of which N rows exist behaves exactly like an [N, C] call -- batch and
    running statistics, every gradient -- and the padding rows come back as zeros."""
    from tests.test_dense_gpu import _modules
    from kpgnn_b200.layers.dense_block import fused_dense_block
    dev = torch.device("cuda:0")
    C = 104
    g = torch.Generator().manual_seed(N)
    x0 = torch.randn(cap, C, generator=g).to(dev)
    r0 = torch.randn(cap, C, generator=g).to(dev)
    gy = torch.randn(cap, C, generator=g).to(dev)
    res = []
    for padded in (False, True):
        lin1, bn1, lin2, bn2, bn3 = _modules(C, C, 3, dev)
        n = cap if padded else N
        x = x0[:n].clone().requires_grad_(True)
        r = r0[:n].clone().requires_grad_(True)
        n_dev = torch.tensor([N], dtype=torch.int32, device=dev) if padded else None
        y = fused_dense_block(x, lin1, bn1, lin2, bn2, bn3, r, n_dev=n_dev)
        gyy = gy[:n].clone()
        if padded:
            gyy[N:] = 0          # upstream never sends gradient into padding rows (pooling ignores them)
        y.backward(gyy)
        out = [y.detach(), x.grad, r.grad] + [p.grad for m in (lin1, bn1, lin2, bn2, bn3) for p in m.parameters()] + \
              [b for m in (bn1, bn2, bn3) for b in (m.running_mean, m.running_var)]
        res.append(out)
    for i, (a, b) in enumerate(zip(*res)):
        if i < 3:
            assert float(b[N:].abs().max()) == 0.0 if b.size(0) > N and i < 2 else True
            b = b[:N]
        assert rel_err(b, a) < 2e-6, (i, rel_err(b, a))


def test_captured_step_over_distinct_batches_matches_oracle(lib):
    """20 optimisation steps over 6 distinct batches of different node / edge counts through ONE captured CUDA graph
    (static padded buffers, device-side row count) vs the oracle model + torch.optim.Adam on the unpadded batches."""
    from kpgnn_b200.model import zinc_kpginplus
    from kpgnn_b200.train import Trainer, fit_spec
    from oracle.model_torch import l1_loss as ol1, zinc_oracle_model
    dev = torch.device("cuda:0")
    G = 24
    hbs = [_host_batch(G, 100 + s) for s in range(6)]
    sizes = {(int(b.x.size(0)), int(b.edge_index.size(1))) for b in hbs}
    assert len(sizes) == 6
    spec, bounds = fit_spec(hbs, 8, 3, 6)
    flats = [spec.pack(b, spec.host_buffer()) for b in hbs]
    torch.manual_seed(0)
    model = zinc_kpginplus(8, 8, 104).to(dev).train()
    ora = zinc_oracle_model(8, 8, 104).to(dev).train()
    ora.load_state_dict(model.state_dict())
    opt = torch.optim.Adam(ora.parameters(), lr=1e-3)
    tr = Trainer(model, spec, bounds, dev)
    # capture() warms up with 3 eager steps + the capture itself does not execute: undo the warm-up's parameter updates
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tr.capture(flats[0])
    model.load_state_dict(sd)
    tr.opt.state.zero_()
    tr.opt.m.zero_()
    tr.opt.v.zero_()
    tr.prefetch(flats[0])
    mine, ref = [], []
    for step in range(20):
        hb = hbs[step % 6]
        mine.append(tr.step_e2e(flats[(step + 1) % 6]))
        ob = {f: getattr(hb, f).to(dev) for f in hb.FIELDS}
        ob["num_graphs"] = G
        opt.zero_grad()
        loss = ol1(ora(ob), ob["y"])
        loss.backward()
        opt.step()
        ref.append(float(loss))
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3), (i, a, b, mine, ref)
    # and the trained parameters agree
    osd = ora.state_dict()
    worst = max(rel_err(v, osd[k]) for k, v in model.state_dict().items() if v.dtype == torch.float32 and v.numel() > 1)
    assert worst < 2e-3, worst
