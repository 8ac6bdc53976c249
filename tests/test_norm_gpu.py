"""FusedBatchNorm1d (csrc/bn.cu) against torch.nn.BatchNorm1d (+ReLU): forward, input/affine gradients, running
statistics, eval-mode and large-batch fallbacks, state_dict compatibility."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


# N=2 is degenerate for the input gradient (analytically zero: both sides are rounding noise), start at 3
@pytest.mark.parametrize("N", [3, 5, 300, 2952, 4096, 4097])
@pytest.mark.parametrize("C", [104, 48])
@pytest.mark.parametrize("relu", [False, True])
def test_bn_matches_torch(lib, N, C, relu):
    from kpgnn_b200.layers.norm import FusedBatchNorm1d
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 7 + C)
    mine = FusedBatchNorm1d(C, relu=relu).to(dev)
    ref = nn.BatchNorm1d(C).to(dev)
    with torch.no_grad():
        mine.weight.copy_(torch.randn(C, generator=g))
        mine.bias.copy_(torch.randn(C, generator=g))
    ref.load_state_dict(mine.state_dict())          # same keys: weight, bias, running_*, num_batches_tracked
    x0 = (torch.randn(N, C, generator=g) * 3 + 1).to(dev)
    gy = torch.randn(N, C, generator=g).to(dev)
    outs = []
    for m, post in ((ref, torch.relu if relu else (lambda t: t)), (mine, lambda t: t)):
        m.train()
        x = x0.clone().requires_grad_(True)
        for _ in range(2):                          # two steps: running statistics accumulate
            y = post(m(x))
        y.backward(gy)
        outs.append((y.detach(), x.grad, m.weight.grad, m.bias.grad, m.running_mean.clone(), m.running_var.clone(),
                     int(m.num_batches_tracked)))
    names = ["y", "dx", "dgamma", "dbeta", "running_mean", "running_var"]
    for n, a, b in zip(names, outs[0][:6], outs[1][:6]):
        scale = max(float(a.abs().max()), 1e-6)
        assert float((a - b).abs().max()) / scale < 2e-5, (n, float((a - b).abs().max()), scale)
    assert outs[0][6] == outs[1][6] == 2
    mine.eval(), ref.eval()
    ye = torch.relu(ref(x0)) if relu else ref(x0)
    assert float((mine(x0) - ye).abs().max()) < 1e-4


def test_bn_deterministic(lib):
    from kpgnn_b200.layers.norm import FusedBatchNorm1d
    dev = torch.device("cuda:0")
    m = FusedBatchNorm1d(104, relu=True).to(dev).train()
    x = torch.randn(3000, 104, device=dev)
    res = []
    for _ in range(2):
        xx = x.clone().requires_grad_(True)
        y = m(xx)
        y.square().sum().backward()
        res.append((y.detach().clone(), xx.grad.clone(), m.weight.grad.clone()))
        m.zero_grad()
    for a, b in zip(*res):
        assert torch.equal(a, b)
