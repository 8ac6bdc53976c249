"""Kineto kernel-time breakdown of extraction + plan build (exploration helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kpgnn_b200 import synth  # noqa: E402
from kpgnn_b200 import data_utils as DU  # noqa: E402
from kpgnn_b200.plan import get_plan  # noqa: E402

dev = torch.device("cuda:0")
cases = [("zinc128_K8", synth.zinc_like_graphs(128, 0), (8, 50, 6, 3, 50, 50, "spd")),
         ("regular1280x16_K6", [synth.regular_graph(1280, 3, s) for s in range(16)], (6, 10, 1, 1, 1, 1, "spd"))]
for name, graphs, args in cases:
    csr = DU.pack_csr(graphs)
    for _ in range(2):
        out = DU._extract_device(csr, *args, device=dev)
        get_plan(out["edge_index"].clone(), out["edge_attr"], csr["N"])
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        out = DU._extract_device(csr, *args, device=dev)
        get_plan(out["edge_index"].clone(), out["edge_attr"], csr["N"])
        torch.cuda.synchronize()
    print("=====", name)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
