"""Timing of the hot-path kernels on the other BASELINE.json configs (parity-test cases; not bench lines).
Run on the GPU box:  python profiles/measure_configs.py > gpurun_out/configs.json"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpgnn_b200 import synth  # noqa: E402
from kpgnn_b200.data_utils import extract_batch  # noqa: E402
from kpgnn_b200.ops import khop_aggregate, ACT_NONE  # noqa: E402
from kpgnn_b200.plan import get_plan  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=10, warm=3):
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    ts = []
    for i in range(reps + warm):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return statistics.mean(ts)


def extraction(name, graphs, args):
    extract_batch(graphs, args, dev)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        b = extract_batch(graphs, args, dev)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    return {"config": name, "graphs": len(graphs), "nodes": b.num_nodes, "E_K": int(b.edge_index.size(1)),
            "ms_per_batch_incl_host_csr_pack": round(dt * 1e3, 3), "graphs_per_s": round(len(graphs) / dt, 1)}, b


def aggregation(name, b, K, d, tables, eps):
    N = b.num_nodes
    ei, ea = b.edge_index, b.edge_attr
    plan, k = get_plan(ei, ea, N)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(N, K, d, device=dev, generator=g, requires_grad=True)
    t0 = torch.randn(5, d, device=dev, generator=g, requires_grad=True) if tables else None
    tk = torch.randn(52, d, device=dev, generator=g, requires_grad=True) if tables else None
    e = torch.zeros(1, device=dev) if eps else None
    out = {}

    def fwd():
        out["y"] = khop_aggregate(x, plan, k, T0=t0, Tk=tk, eps=e, act=ACT_NONE)
    ms_f = timed(fwd)
    gy = torch.randn_like(out["y"])

    def fb():
        y = khop_aggregate(x, plan, k, T0=t0, Tk=tk, eps=e, act=ACT_NONE)
        y.backward(gy)
    ms_fb = timed(fb)
    alg = 4 * N * K * d * 2 + 4 * (N * K + 1) + plan.nnz * (6 if tables else 4)
    return {"config": name, "N": N, "nnz": plan.nnz, "k": K, "d": d, "fwd_ms": round(ms_f, 4),
            "fwd_plus_bwd_ms": round(ms_fb, 4), "fwd_algorithmic_bytes": alg,
            "fwd_GBps": round(alg / ms_f / 1e6, 1)}


res = []
r, b = extraction("zinc128_K8_spd", synth.zinc_like_graphs(128, 0), (8, 50, 6, 3, 50, 50, "spd"))
res.append(r)
r, b16 = extraction("zinc128_K16_spd", synth.zinc_like_graphs(128, 0), (16, 50, 6, 3, 50, 50, "spd"))
res.append(r)
res.append(aggregation("prime_layer1_K16_dk6 (generic float2 path)", b16, 16, 6, True, True))
r, bz = extraction("zinc2048_K8_spd", synth.zinc_like_graphs(2048, 1), (8, 50, 6, 3, 50, 50, "spd"))
res.append(r)
r, breg = extraction("regular1280_x8_K6_spd", [synth.regular_graph(1280, 3, s) for s in range(8)], (6, 10, 1, 1, 1, 1, "spd"))
res.append(r)
res.append(aggregation("regular1280_x8_KGIN_K6_d16 (no tables, self term)", breg, 6, 16, False, True))
import networkx as nx
sr = []
for s in range(15):
    G = nx.random_regular_graph(12, 25, seed=s)
    e = np.array(list(G.to_directed().edges)).T
    e = e[:, np.lexsort((e[1], e[0]))]
    sr.append({"num_nodes": 25, "x": np.ones(25, dtype=np.int64), "edge_index": e.astype(np.int64), "edge_attr": None})
r, bsr = extraction("sr25shape_x15_K4_gd", sr, (4, 1000, 4, 1, 1000, 1000, "gd"))
res.append(r)
print(json.dumps(res, indent=1))
