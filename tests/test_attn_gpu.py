"""AttentionCombine kernels (kp_attn_combine_*, csrc/attn.cu) against the reference formulation -- torch.nn.LSTM
(bidirectional, hidden size K) + sum + softmax + weighted sum, layers/combine.py:8-27 -- evaluated in float64 on the
same device: output, dX and all eight LSTM parameter gradients, at every (K, d) the BASELINE configs produce
(KP-GIN+ ZINC 8 x 104; EXP 3 x 16; SR25 4 x 12; KPGINPrime 16 x 6; regular-graph 6 x 16) plus the limits."""
import pytest
import torch

from oracle import layers_torch as OL
from tests.util import RTOL, rel_err

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("K,d", [(8, 104), (3, 16), (4, 12), (16, 6), (6, 16), (16, 128), (2, 5), (1, 24), (5, 77),
                                 (12, 64)])
@pytest.mark.parametrize("N", [1, 37, 3000])
def test_attention_combine_matches_lstm_formulation(lib, K, d, N):
    from kpgnn_b200.layers.combine import AttentionCombine
    dev = torch.device("cuda:0")
    torch.manual_seed(K * 1000 + d)
    mine = AttentionCombine(d, K).to(dev)
    # default init is U(-1/sqrt(K), 1/sqrt(K)); scale up so the gates leave their linear range
    with torch.no_grad():
        for p in mine.parameters():
            p.mul_(2.0)
    ora = OL.OracleAttentionCombine(d, K).to(dev).double()
    ora.load_state_dict({k: v.double() for k, v in mine.state_dict().items()})
    x0 = torch.randn(N, K, d, device=dev)
    gy = torch.randn(N, d, device=dev)
    x = x0.clone().requires_grad_(True)
    y = mine(x)
    y.backward(gy)
    xr = x0.double().requires_grad_(True)
    yr = ora(xr)
    yr.backward(gy.double())
    assert rel_err(y, yr) < RTOL, ("out", rel_err(y, yr))
    assert rel_err(x.grad, xr.grad) < RTOL, ("dx", rel_err(x.grad, xr.grad))
    gr = dict(ora.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in gr.values())
    for n, p in mine.named_parameters():
        err = rel_err(p.grad, gr[n].grad, floor=1e-3 * gmax)
        assert err < RTOL, (n, err)


def test_attention_combine_strided_input_and_determinism(lib):
    """x as a hop-sliced view of a wider tensor (what KPGINPlus hands over for k < K) and bitwise reproducibility."""
    from kpgnn_b200.layers.combine import AttentionCombine
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    K, d, N = 5, 104, 777
    mine = AttentionCombine(d, K).to(dev)
    big = torch.randn(N, 8, d, device=dev)
    res = []
    for view in (big[:, 1:1 + K], big[:, 1:1 + K].contiguous(), big[:, 1:1 + K]):
        x = view.detach().requires_grad_(True)
        for p in mine.parameters():
            p.grad = None
        y = mine(x)
        y.square().sum().backward()
        res.append([y.detach().clone(), x.grad.clone()] + [p.grad.clone() for p in mine.parameters()])
    for a, b, c in zip(*res):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_attention_combine_large_batch(lib):
    """Persistent loop (more nodes than resident warps) at the roofline batch's node count."""
    from kpgnn_b200.layers.combine import AttentionCombine
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    K, d, N = 8, 104, 189489
    mine = AttentionCombine(d, K).to(dev)
    ora = OL.OracleAttentionCombine(d, K).to(dev)
    ora.load_state_dict(mine.state_dict())
    x0 = torch.randn(N, K, d, device=dev)
    x = x0.clone().requires_grad_(True)
    y = mine(x)
    y.sum().backward()
    xr = x0.clone().requires_grad_(True)
    yr = ora(xr)
    yr.sum().backward()
    assert rel_err(y, yr) < RTOL and rel_err(x.grad, xr.grad) < RTOL
    gr = dict(ora.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in gr.values())
    for n, p in mine.named_parameters():
        # sums over 1.5 M (node, hop) terms in fp32 on both sides (cuDNN vs fixed-order partials)
        assert rel_err(p.grad, gr[n].grad, floor=1e-2 * gmax) < 5 * RTOL, n
