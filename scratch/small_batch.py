"""Small-batch (L2-resident) regime of the aggregation kernels: the bench batch (128 graphs) with a warm L2, as
inside the training step.  usage: python scratch/small_batch.py [GRAPHS] [--once] [--k K ...]
Prints per-launch microseconds (CUDA events around REP back-to-back launches, so launch gaps are included).
--once: one launch of forward and backward per k (for ncu)."""
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kpgnn_b200 import _lib
from kpgnn_b200.ops import _make_desc, ACT_GELU, ACT_NONE
from kpgnn_b200.plan import get_plan

args = [a for a in sys.argv[1:] if not a.startswith("--")]
once = "--once" in sys.argv
G = int(args[0]) if args else 128
ks = [int(a) for a in args[1:]] or [1, 2, 4, 8]
REP = 1 if once else 40
dev = torch.device("cuda:0")
hb = bench.host_batch(G, seed=0)
ei, ea = hb.edge_index.to(dev), hb.edge_attr.to(dev)
N = hb.x.size(0)
K, H = bench.K, bench.HIDDEN
gen = torch.Generator(device=dev).manual_seed(0)
# X as in the training step: hop slices of the layer-history buffer [N, L+1, H] (strided view)
hist = torch.randn(N, K + 1, H, device=dev, generator=gen)
Pfull = torch.randn(N, K, H, device=dev, generator=gen)
t0 = torch.randn(5, H, device=dev, generator=gen)
tk = torch.randn(52, H, device=dev, generator=gen)
lib = _lib.lib()
st = torch.cuda.current_stream(dev)
sp = C.c_void_p(st.cuda_stream)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(1 if once else 7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(REP):
            fn()
        b.record(st)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / REP)
    return statistics.median(ts), min(ts)


print("graphs %d  nodes %d  lib %s" % (G, N, os.environ.get("KPGNN_B200_LIB", "default")), flush=True)
plan, _ = get_plan(ei, ea, N)
for k in ks:
    x = hist[:, :k, :]
    P = Pfull[:, :k, :]
    th = torch.softmax(torch.randn(k, H, device=dev, generator=gen), 0)
    out = torch.empty(N, H, device=dev)

    def cfg(name, P_, t0_, tk_, act, fuse):
        o = out if fuse else torch.empty(N, k, H, device=dev)
        desc = _make_desc(plan, k, x, P_, t0_, tk_, th if fuse else None, None, act, fuse, False, False)
        med, mn = timed(lambda: _lib.check(lib.kp_agg_forward(C.byref(desc), o.data_ptr(), sp), "fwd"))
        print("k=%d fwd %-28s %7.2f us (min %.2f)" % (k, name, med, mn), flush=True)
        return desc

    desc = cfg("full (GELU,P,tables,fuse)", P, t0, tk, ACT_GELU, True)
    if not once:
        cfg("no P", None, t0, tk, ACT_GELU, True)
        cfg("no tables", P, None, None, ACT_GELU, True)
        cfg("gather only (B2 shape)", None, None, None, ACT_NONE, False)
    dout = torch.randn(N, H, device=dev, generator=gen)
    dX = torch.empty(N, k, H, device=dev)
    dP = torch.empty(N, k, H, device=dev)
    dT0, dTk, dth = torch.empty_like(t0), torch.empty_like(tk), torch.empty_like(th)
    nb = C.c_size_t(0)
    _lib.check(lib.kp_agg_backward_workspace_bytes(C.byref(desc), C.byref(nb)), "ws")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    med, mn = timed(lambda: _lib.check(lib.kp_agg_backward(
        C.byref(desc), dout.data_ptr(), dX.data_ptr(), dP.data_ptr(), dT0.data_ptr(), dTk.data_ptr(),
        dth.data_ptr(), None, ws.data_ptr(), ws.numel(), sp), "bwd"))
    print("k=%d bwd B1+B2+B3+reductions            %7.2f us (min %.2f)" % (k, med, mn), flush=True)
