"""FusedAdam (csrc/adam.cu) against torch.optim.Adam with the reference's settings (train_ZINC.py:244)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch(lib):
    from kpgnn_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shapes = [(1,), (7,), (104,), (104, 104), (5000,), (52, 104), (3, 1025)]
    pa = [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    a, b = FusedAdam(pa, lr=1e-3), torch.optim.Adam(pb, lr=1e-3)
    for step in range(6):
        for i, (x, y) in enumerate(zip(pa, pb)):
            if i == 1 and step < 2:
                x.grad = y.grad = None                      # a parameter without a gradient is skipped
                continue
            gr = (torch.randn(x.shape, generator=g) * (10.0 ** (i - 3))).to(dev)
            x.grad, y.grad = gr.clone(), gr.clone()
        a.step()
        b.step()
    for i, (x, y) in enumerate(zip(pa, pb)):
        if i == 1:
            continue        # torch keeps a per-parameter step count; ours is global (all parameters always have grads in training)
        assert float((x - y).abs().max()) <= 2e-6 * max(float(y.abs().max()), 1.0), i
        st = b.state[y]
        assert float((a.exp_avg(i) - st["exp_avg"]).abs().max()) <= 1e-6 * max(float(st["exp_avg"].abs().max()), 1e-12)
        assert float((a.exp_avg_sq(i) - st["exp_avg_sq"]).abs().max()) <= 1e-6 * max(float(st["exp_avg_sq"].abs().max()), 1e-12)


def test_fused_adam_in_cuda_graph(lib):
    from kpgnn_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    p = torch.ones(3000, device=dev, requires_grad=True)
    q = p.detach().clone().requires_grad_(True)
    a, b = FusedAdam([p], lr=1e-2), torch.optim.Adam([q], lr=1e-2)
    gbuf = torch.zeros(3000, device=dev)
    p.grad = gbuf
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a.step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        a.step()
    q.grad = torch.zeros(3000, device=dev)
    b.step()                                                # the captured step did not execute, the warm-up one did
    for i in range(4):
        gbuf.fill_(float(i + 1))
        graph.replay()
        q.grad = torch.full((3000,), float(i + 1), device=dev)
        b.step()
    torch.cuda.synchronize()
    assert float((p - q).abs().max()) < 1e-5
