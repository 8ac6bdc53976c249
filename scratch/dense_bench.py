"""Times the fused dense block alone (CUDA events, warm) -- tuning helper, not a bench value."""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpgnn_b200.layers.dense_block import fused_dense_block
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, C = int(os.environ.get("N", 2952)), 104
lin1, bn1, lin2, bn2, bn3 = (nn.Linear(C, C), nn.BatchNorm1d(C), nn.Linear(C, C), nn.BatchNorm1d(C), nn.BatchNorm1d(C))
mods = [m.to(dev).train() for m in (lin1, bn1, lin2, bn2, bn3)]
x = torch.randn(N, C, device=dev, requires_grad=True)
r = torch.randn(N, C, device=dev, requires_grad=True)
gy = torch.randn(N, C, device=dev)
def fwd():
    return fused_dense_block(x, *mods, r)
for _ in range(5):
    y = fwd(); y.backward(gy)
torch.cuda.synchronize()
def timeit(fn, n=50):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
tf = timeit(lambda: fwd())
def fb():
    y = fwd(); y.backward(gy)
tfb = timeit(fb)
print("rows/CTA env=%s  N=%d: fwd %.1f us, fwd+bwd %.1f us (bwd %.1f us)" % (os.environ.get("KP_DENSE_ROWS", "default"), N, tf, tfb, tfb - tf))
