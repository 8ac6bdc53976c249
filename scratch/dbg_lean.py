import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpgnn_b200 import _lib
from kpgnn_b200.ops import khop_aggregate, ACT_GELU, ACT_NONE
from kpgnn_b200.plan import get_plan
from tests.util import zinc_batch
dev = torch.device("cuda:0")
lib = _lib.lib()
for K, H, ng in ((1, 104, 6), (3, 104, 6), (8, 104, 6), (8, 104, 300), (4, 64, 50)):
    b = zinc_batch(ng, K, "spd", seed=K)
    N = b["num_nodes"]
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    plan, k = get_plan(ei, ea, N)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(N, K, H, device=dev, generator=g)
    P = torch.randn(N, K, H, device=dev, generator=g)
    t0 = torch.randn(5, H, device=dev, generator=g)
    tk = torch.randn(52, H, device=dev, generator=g)
    th = torch.softmax(torch.randn(K, H, device=dev, generator=g), 0)
    for fuse in (True, False):
        for act in (ACT_GELU, ACT_NONE):
            if act == ACT_NONE and fuse:
                continue
            for useP in (True, False):
                outs = []
                for flag in (16, 4, 1):
                    lib.kp_agg_set_force_generic(flag)
                    y = khop_aggregate(x, plan, k, P=P if useP else None, T0=t0, Tk=tk, theta=th if fuse else None,
                                       act=act, fuse=fuse)
                    torch.cuda.synchronize()
                    outs.append(y)
                lib.kp_agg_set_force_generic(0)
                d1 = (outs[0] - outs[2]).abs().max().item()
                d2 = (outs[1] - outs[2]).abs().max().item()
                bad = (outs[0] - outs[2]).abs().amax(dim=tuple(range(1, outs[0].dim())))
                nb = int((bad > 1e-4).sum())
                print(f"K={K} H={H} N={N} fuse={fuse} act={act} P={useP}: lean-generic {d1:.3e} ring-generic {d2:.3e} bad nodes {nb}",
                      (bad > 1e-4).nonzero().flatten()[:10].tolist())
