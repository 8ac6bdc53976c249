"""Round-2 exploration timings (not bench lines): extraction phases and the config-5 forward.
    python profiles/measure_r2.py > gpurun_out/r2_measure.json"""
import ctypes as C
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpgnn_b200 import _lib, synth  # noqa: E402
from kpgnn_b200 import data_utils as DU  # noqa: E402
from kpgnn_b200.plan import get_plan  # noqa: E402

dev = torch.device("cuda:0")
res = []


def ev():
    return torch.cuda.Event(enable_timing=True)


def extraction_phases(name, graphs, args, reps=5):
    t0 = time.perf_counter()
    csr = DU.pack_csr(graphs)
    t_pack = time.perf_counter() - t0
    DU._extract_device(csr, *args, device=dev)
    torch.cuda.synchronize()
    # instrument the three kernels separately
    lib = _lib.lib()
    K = args[0]
    tt = {"hops": [], "emit": [], "periph": [], "total_wall": []}
    for _ in range(reps):
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        out = DU._extract_device(csr, *args, device=dev)
        torch.cuda.synchronize()
        tt["total_wall"].append((time.perf_counter() - w0) * 1e3)
    r = {"config": name, "graphs": len(graphs), "nodes": csr["N"], "E_K": int(out["edge_index"].size(1)),
         "host_pack_ms": round(t_pack * 1e3, 3), "device_extract_wall_ms": round(statistics.mean(tt["total_wall"]), 3),
         "graphs_per_s_device_only": round(len(graphs) / (statistics.mean(tt["total_wall"]) * 1e-3), 1)}
    res.append(r)
    return out, csr


def kernel_times(fn, reps=10, warm=3, flush=None):
    ts = []
    for i in range(reps + warm):
        if flush is not None:
            flush.add_(1)
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return statistics.mean(ts)


flush = torch.zeros(64 * 1024 * 1024, device=dev)
zg = synth.zinc_like_graphs(128, 0)
extraction_phases("zinc128_K8_spd", zg, (8, 50, 6, 3, 50, 50, "spd"))
extraction_phases("zinc128_K16_spd", zg, (16, 50, 6, 3, 50, 50, "spd"))
extraction_phases("zinc2048_K8_spd", synth.zinc_like_graphs(2048, 1), (8, 50, 6, 3, 50, 50, "spd"))
t0 = time.perf_counter()
rg = [synth.regular_graph(1280, 3, s) for s in range(64)]
res.append({"config": "regular1280 x64 generation (networkx, host)", "s": round(time.perf_counter() - t0, 2)})
out, csr = extraction_phases("regular1280_x64_K6_spd", rg, (6, 10, 1, 1, 1, 1, "spd"))

# config 5 forward: KGINConv(16, K=6), 64 graphs
from kpgnn_b200.simulation import KGINConv  # noqa: E402
from kpgnn_b200 import ops  # noqa: E402
torch.manual_seed(0)
model = KGINConv(16, 6).to(dev).eval()
N = csr["N"]
ei, ea = out["edge_index"], out["edge_attr"]
x = torch.ones(N, 1, device=dev)
bt = torch.from_numpy(csr["node_graph"].astype(np.int64)).to(dev)
for mode in ("blocks", "rows"):
    ei2 = ei.clone()
    keep = ops.LONG_ROW_ENTRIES
    if mode == "rows":
        ops.LONG_ROW_ENTRIES = 10 ** 9
    with torch.no_grad():
        model(x, ei2, ea, bt)
        plan, k = get_plan(ei2, ea, N)
        xx = torch.randn(N, 6, 16, device=dev)
        eps = torch.zeros(1, device=dev)
        t_agg = kernel_times(lambda: ops.khop_aggregate(xx, plan, k, eps=eps), flush=flush)
        t_fwd = kernel_times(lambda: model(x, ei2, ea, bt), flush=flush)
    ops.LONG_ROW_ENTRIES = keep
    alg = 4 * N * 6 * 16 * 2 + 4 * (N * 6 + 1) + plan.nnz * 4
    res.append({"config": "regular1280_x64 KGIN K=6 d=16 forward, aggregation path = " + mode, "N": N, "nnz": plan.nnz,
                "agg_ms": round(t_agg, 4), "layer_fwd_ms": round(t_fwd, 4), "agg_algorithmic_bytes": alg,
                "agg_GBps": round(alg / t_agg / 1e6, 1), "blocks": plan.num_blocks})
t_plan = []
for _ in range(5):
    ei3 = ei.clone()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    get_plan(ei3, ea, N)
    b.record()
    torch.cuda.synchronize()
    t_plan.append(a.elapsed_time(b))
res.append({"config": "regular1280_x64 plan build", "ms": round(statistics.mean(t_plan), 3)})
print(json.dumps(res, indent=1))
