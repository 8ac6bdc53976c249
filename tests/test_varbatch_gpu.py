"""Batches of different sizes through static buffers: the compact wire format and its device-side
unpacking, the dense block's device-side row count, and a captured training step replayed over distinct batches --
loss trajectory against the oracle model stepping on the same (unpadded) batches with torch.optim.Adam."""
import copy

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
        if i in (4, 8):
            # lin1.bias / lin2.bias feed a BatchNorm: their gradient is analytically zero, both sides are rounding noise
            # whose pattern depends on the slab decomposition (capacity vs exact row count) -- hold it to the weight scale
            w = res[0][i - 1]
            assert float((a - b).abs().max()) < 2e-6 * float(w.abs().max()), (i, float((a - b).abs().max()))
            continue
        assert rel_err(b, a) < 2e-6, (i, rel_err(b, a))


def test_captured_step_over_distinct_batches_matches_oracle(lib):
    """ONE captured CUDA graph (static padded buffers, device-side row count, compact wire upload, loss read back every
    step) serving 6 distinct batches of different node / edge counts, next to the oracle model + torch.optim.Adam on the
    unpadded batches.
    (a) Every batch shape from the same initial state: loss within 1e-5 and every parameter gradient within 1e-4 of the
        oracle's -- the padding rows / masked edges / device-side row count change nothing.
    (b) 20 free-running optimisation steps cycling through the batches: the loss trajectories agree to 1e-4 until the
        training dynamics amplify rounding differences (Adam's first updates are sign(g) * lr, the L1 loss has gradient
        sign(score - y): a parameter whose gradient is rounding noise, or a prediction crossing its target, moves two
        fp32 runs apart whatever their accuracy -- measured with Adam's eps raised to 1e-2: 1e-7 agreement for 8 steps,
        then 1.5e-4 in one step when a prediction crossed its target; with the default eps: 1e-7, 2e-5, 4e-4, 7e-3 on
        steps 0..3, 16 % by step 13), so the run is held to 1e-4 for its first 2 steps; for the other 18 only that both
        runs stay finite, reduce the loss, and have mean losses within 25 % of each other."""
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
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("alphas"):
                p.add_(0.3 * torch.randn_like(p))
    ora = zinc_oracle_model(8, 8, 104).to(dev).train()
    tr = Trainer(model, spec, bounds, dev)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    tr.capture(flats[0])            # warms up with eager steps (the capture itself executes nothing)
    assert tr.graph is not None

    def reset():
        model.load_state_dict(sd)
        ora.load_state_dict(sd)
        tr.opt.state.zero_()
        tr.opt.m.zero_()
        tr.opt.v.zero_()

    def oracle_batch(hb):
        ob = {f: getattr(hb, f).to(dev) for f in hb.FIELDS}
        ob["num_graphs"] = G
        return ob
    oparams = dict(ora.named_parameters())
    # ---- (a)
    # The product path is bit-reproducible; the ORACLE on a GPU is not: its index_add_ sums are atomically ordered, and a
    # batch that happens to hold an activation within rounding of a ReLU / |.| kink then yields one of two discrete
    # gradients from run to run (observed on 2 of these 6 batches: the oracle differed from ITSELF by 2.7e-2 between
    # runs while 48 replays of the product were bitwise identical).  So the oracle is evaluated up to 4 times per batch
    # and the product has to match one of its outcomes.
    for i, hb in enumerate(hbs):
        reset()
        tr.prefetch(flats[i])
        mine = tr.step_e2e(flats[i])
        mygrads = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
        ob = oracle_batch(hb)
        worst = None
        for attempt in range(4):
            for p in ora.parameters():
                p.grad = None
            loss = ol1(ora(ob), ob["y"])
            loss.backward()
            ref = float(loss.detach())
            assert abs(mine - ref) <= 1e-5 * max(abs(ref), 1e-3), (i, mine, ref)
            gmax = max(float(p.grad.abs().max()) for p in oparams.values() if p.grad is not None)
            worst = None
            for n, g in mygrads.items():
                if n.endswith(("mlp.0.bias", "mlp.3.bias")):
                    continue
                og = oparams[n].grad if oparams[n].grad is not None else torch.zeros_like(g)
                err = rel_err(g, og, floor=1e-2 * gmax)
                # the scalar gates pew / pcw are sums of ~N*K*H signed products that cancel to ~1 % of their running
                # partial sums: fp32 summation order alone moves them by ~1e-3
                if err >= (1e-2 if g.numel() == 1 else 1e-4) and (worst is None or err > worst[1]):
                    worst = (n, err)
            if worst is None:
                break
        assert worst is None, (i, worst)
    # ---- (b)
    reset()
    opt = torch.optim.Adam(ora.parameters(), lr=1e-3)
    tr.prefetch(flats[0])
    mine_all, ref_all = [], []
    for step in range(20):
        mine = tr.step_e2e(flats[(step + 1) % 6])
        ob = oracle_batch(hbs[step % 6])
        opt.zero_grad()
        loss = ol1(ora(ob), ob["y"])
        loss.backward()
        opt.step()
        ref = float(loss.detach())
        mine_all.append(mine)
        ref_all.append(ref)
        if step < 2:
            assert abs(mine - ref) <= 1e-4 * max(abs(ref), 1e-3), (step, mine, ref)
    assert all(np.isfinite(mine_all))
    # both runs train: the mean loss of the last 6 steps (one pass over the batches) is below that of the first 6, and
    # the two runs' means over the 20 steps are within 25 % of each other
    assert np.mean(mine_all[-6:]) < np.mean(mine_all[:6]) and np.mean(ref_all[-6:]) < np.mean(ref_all[:6])
    assert abs(np.mean(mine_all) - np.mean(ref_all)) < 0.25 * np.mean(ref_all), (mine_all, ref_all)


def test_pipelined_e2e_steps_equal_synchronous_steps(lib):
    """Trainer.step_e2e_pipelined (loss of step i read while step i+1 runs, sticky plan statistics checked one step late)
    must produce the loss sequence of the synchronous step_e2e bit for bit, and still catch a batch that overflows the
    plan capacity."""
    from kpgnn_b200.model import zinc_kpginplus
    from kpgnn_b200.train import Trainer, fit_spec
    dev = torch.device("cuda:0")
    hbs = [_host_batch(16, 300 + s) for s in range(4)]
    spec, bounds = fit_spec(hbs, 8, 3, 6)
    flats = [spec.pack(b, spec.host_buffer()) for b in hbs]
    seqs = []
    for pipelined in (False, True):
        torch.manual_seed(0)
        model = zinc_kpginplus(8, 8, 64).to(dev).train()
        tr = Trainer(model, spec, bounds, dev)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        tr.capture(flats[0])
        model.load_state_dict(sd)
        tr.opt.state.zero_()
        tr.opt.m.zero_()
        tr.opt.v.zero_()
        tr.prefetch(flats[0])
        out = []
        for step in range(10):
            nxt = flats[(step + 1) % 4]
            if pipelined:
                v = tr.step_e2e_pipelined(nxt)
                if v is not None:
                    out.append(v)
            else:
                out.append(tr.step_e2e(nxt))
        if pipelined:
            out.append(tr.drain())
        seqs.append(out)
    assert len(seqs[0]) == len(seqs[1]) == 10
    assert seqs[0] == seqs[1], (seqs[0], seqs[1])
    # overflow: shrink the validated capacity below what the batches need; the lagged check must raise within two steps
    tr.plan_obj.capacity = 8
    tr.prefetch(flats[0])
    with pytest.raises(Exception):
        for step in range(3):
            tr.step_e2e_pipelined(flats[(step + 1) % 4])
        tr.drain()
