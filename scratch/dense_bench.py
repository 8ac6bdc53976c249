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
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        y = fwd(); y.backward(gy)
    torch.cuda.synchronize()
import collections
agg = collections.defaultdict(list)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:60]].append(e.time_range.end - e.time_range.start)
for k, v in agg.items():
    print("%-62s n=%3d  min %.1f  med %.1f us" % (k, len(v), min(v), sorted(v)[len(v) // 2]))
