import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpgnn_b200 import _lib
from kpgnn_b200.ops import khop_aggregate, ACT_GELU, ACT_NONE
from kpgnn_b200.plan import get_plan
from tests.util import zinc_batch
dev = torch.device("cuda:0")
lib = _lib.lib()
K, H, ng = 1, 104, 2
b = zinc_batch(ng, K, "spd", seed=K)
N = b["num_nodes"]
ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
plan, k = get_plan(ei, ea, N)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, K, H, device=dev, generator=g)
t0 = torch.randn(5, H, device=dev, generator=g)
tk = torch.randn(52, H, device=dev, generator=g)
print("rowptr", plan.rowptr[:8].tolist(), "col", plan.col[:8].tolist(), "attr", plan.attr16[:8].tolist())
for tabs in (False, True):
    outs = []
    for flag in (1, 0):
        lib.kp_agg_set_force_generic(flag)
        y = khop_aggregate(x, plan, k, T0=t0 if tabs else None, Tk=tk if tabs else None, act=ACT_NONE, fuse=False)
        torch.cuda.synchronize()
        outs.append(y)
    lib.kp_agg_set_force_generic(0)
    print("tabs", tabs, "maxdiff", (outs[0] - outs[1]).abs().max().item())
    print(" generic node0", outs[0][0, 0, :6].tolist())
    print(" lean    node0", outs[1][0, 0, :6].tolist())
    # what would lean equal? try candidates
    c0 = plan.col[plan.rowptr[0]:plan.rowptr[1]].long()
    print(" cols", c0.tolist(), " sum X[cols]", x[c0, 0, :6].sum(0).tolist())
    if tabs:
        a0 = plan.attr16[plan.rowptr[0]:plan.rowptr[1]].long()
        print(" attrs", a0.tolist(), "sum T0", t0[a0, :6].sum(0).tolist())
