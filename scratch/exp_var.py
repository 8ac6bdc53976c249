import sys, os, statistics, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kpgnn_b200 import _lib
from kpgnn_b200.ops import _make_desc, ACT_GELU, ACT_NONE
from kpgnn_b200.plan import get_plan
dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
hb = bench.host_batch(G, seed=1000 + G)
ei, ea = hb.edge_index.to(dev), hb.edge_attr.to(dev)
N = hb.x.size(0)
plan, k = get_plan(ei, ea, N)
K, H = bench.K, bench.HIDDEN
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, K, H, device=dev, generator=g)
P = torch.randn(N, K, H, device=dev, generator=g)
t0 = torch.randn(5, H, device=dev, generator=g)
tk = torch.randn(52, H, device=dev, generator=g)
th = torch.softmax(torch.randn(K, H, device=dev, generator=g), 0)
lib = _lib.lib()
st = torch.cuda.current_stream(dev)
sp = C.c_void_p(st.cuda_stream)
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def run(name, P_, t0_, tk_, act, fuse):
    out = torch.empty((N, H) if fuse else (N, K, H), device=dev)
    desc = _make_desc(plan, k, x, P_, t0_, tk_, th if fuse else None, None, act, fuse, False, False)
    ts = []
    for i in range(13):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        _lib.check(lib.kp_agg_forward(C.byref(desc), out.data_ptr(), sp), "fwd")
        b.record(st)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    nb = 4 * N * K * H * (2 if P_ is not None else 1) + 4 * N * H * (1 if fuse else K) + 4 * (N * K + 1) + plan.nnz * 6
    ms = statistics.mean(ts)
    print(f"{name:34s} {ms*1e3:8.1f} us  {nb/ms/1e6:8.1f} GB/s  min {min(ts)*1e3:.1f}")
run("full (GELU,P,tables,fuse)", P, t0, tk, ACT_GELU, True)
run("no P", None, t0, tk, ACT_GELU, True)
run("no tables", P, None, None, ACT_GELU, True)
run("no P no tables", None, None, None, ACT_GELU, True)
run("ACT_NONE unfused (writes NKH)", P, t0, tk, ACT_NONE, False)
run("GELU unfused", P, t0, tk, ACT_GELU, False)
run("no P, no tab, ACT_NONE unfused", None, None, None, ACT_NONE, False)
