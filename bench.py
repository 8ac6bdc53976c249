#!/usr/bin/env python
"""bench.py -- ZINC-shape KP-GIN+ (K=8, 8 layers, hidden 104, residual) TRAINING graphs/s on N B200s, plus the
K-hop aggregation kernel's achieved HBM bandwidth against the measured roofline, plus the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one optimisation step (forward + backward + Adam) over one synthetic batch of 128 ZINC-shaped
graphs per GPU, extraction outputs in the reference's wire layout (int64 edge_index [2,E_K], edge_attr [E_K,K],
peripheral attrs).  Ours: the whole step is one CUDA graph; the graph plan is rebuilt from the raw int64 batch
EVERY step (a new batch arrives every step in training) and that is inside every timed region.
  value : inputs resident in HBM  -> plan rebuild + graph replay
  e2e   : inputs in pinned host memory -> H2D copy of the batch + plan rebuild + graph replay + D2H loss
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRAPHS_PER_GPU = 128
K, LAYERS, HIDDEN = 8, 8, 104
EXTRACT_ARGS = (K, 50, 6, 3, 50, 50, "spd")          # train_ZINC.py:124-134
ROOFLINE_GRAPHS = 8192                               # >= 1 GB of algorithmic bytes per forward call
METRIC = "ZINC-shape KPGINPlus K=8 L=8 H=104 training throughput"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------
def host_batch(num_graphs, seed):
    """Synthetic ZINC-shaped batch in the reference wire layout, on the host (untimed setup)."""
    from kpgnn_b200 import synth
    from kpgnn_b200.model import Batch
    from kpgnn_b200.data_utils import extract_batch_host
    graphs = synth.zinc_like_graphs(num_graphs, seed=seed)
    fields = extract_batch_host(graphs, EXTRACT_ARGS)
    fields["y"] = torch.tensor([g["y"] for g in graphs], dtype=torch.float32)
    return Batch(**fields)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
NUM_BATCHES = 8                                      # distinct seeded batches (different N / E / nnz) the timed steps rotate through


class BenchStream(object):
    """The bench's stream of batches and its trainer (kpgnn_b200/train.py): NUM_BATCHES distinct 128-graph batches per
    rank, packed in the compact wire format into pinned host buffers (the loader's side, untimed), one captured step
    graph at the capacity of the largest batch, every timed step on the NEXT batch."""

    def __init__(self, device, rank, world, eager=False):
        from kpgnn_b200.model import zinc_kpginplus
        from kpgnn_b200.train import Trainer, fit_spec
        hbs = [host_batch(GRAPHS_PER_GPU, seed=1000 * rank + s) for s in range(NUM_BATCHES)]
        self.sizes = [(int(b.x.size(0)), int(b.edge_index.size(1))) for b in hbs]
        self.wire_bytes_reference_layout = int(statistics.mean(b.nbytes() for b in hbs))
        self.spec, self.bounds = fit_spec(hbs, K, EXTRACT_ARGS[3], EXTRACT_ARGS[2])
        self.flats = [self.spec.pack(b, self.spec.host_buffer()) for b in hbs]
        self.dev_flats = [f.to(device) for f in self.flats]              # "value": batches resident in HBM
        torch.manual_seed(0)
        model = zinc_kpginplus(K, LAYERS, HIDDEN).to(device).train()
        self.tr = Trainer(model, self.spec, self.bounds, device, world=world, lr=1e-3, use_graph=not eager)
        self.tr.capture(self.flats[0])
        self.i = 0
        self.tr.prefetch(self.flats[0])

    def step_resident(self):
        """Next batch already in HBM: device-to-device hand-over into the staging buffer + the step graph."""
        self.i = (self.i + 1) % NUM_BATCHES
        self.tr.wire.stage.copy_(self.dev_flats[self.i], non_blocking=True)
        self.tr.replay()

    def step_e2e(self):
        """Next batch in pinned host memory: its upload was issued behind the previous step; hand-over, step graph,
        upload of the batch after it, loss read-back, deferred plan validation."""
        self.i = (self.i + 1) % NUM_BATCHES
        return self.tr.step_e2e_pipelined(self.flats[(self.i + 1) % NUM_BATCHES])


def flush_l2(buf):
    buf.add_(1)


def timed_steps(fn, steps, device, flush_buf, dist_on):
    """EXACTLY `steps` steps bracketed by a barrier + synchronize on both sides; inside the bracket every step has its own
    pair of CUDA events on the launching stream and the L2 flush runs between one step's end event and the next step's
    start event (in stream order, so it is outside every timed interval).  The host does not synchronise between steps
    (a training loop does not either): launches queue behind the flush, ranks stay coupled only through the step's own
    gradient exchange.  Returns the per-step device times in ms."""
    st = torch.cuda.current_stream(device)
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    evs = []
    for _ in range(steps):
        flush_l2(flush_buf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        evs.append((a, b))
    torch.cuda.synchronize(device)
    if dist_on:
        torch.distributed.barrier()
    return [a.elapsed_time(b) for a, b in evs]


def _check_agg_launch(hb, ei, ea, x, P, t0, tk, th, out, dout=None, dX=None, graphs=48):
    """The timed launch is CHECKED before it is timed: its output rows (and, when given, its dX rows) for the first and
    the last `graphs` graphs of the batch against the textbook formula evaluated in float64 with plain torch ops
    (masked gather + embedding lookup, index_add at the destination, GELU, + P, theta-weighted sum over hops:
    layers/KPGINplus.py:74-88, combine.py:43-58).  Graphs are independent, so a chunk is checked on its own."""
    batch = hb.batch
    G = int(batch[-1]) + 1
    worst = 0.0
    for g0, g1 in ((0, min(graphs, G)), (max(0, G - graphs), G)):
        n0 = int(torch.searchsorted(batch, torch.tensor(g0)))
        n1 = int(torch.searchsorted(batch, torch.tensor(g1)))
        src_all = hb.edge_index[0]
        e0 = int(torch.searchsorted(src_all, torch.tensor(n0)))
        e1 = int(torch.searchsorted(src_all, torch.tensor(n1)))
        src, dst, a = ei[0, e0:e1] - n0, ei[1, e0:e1] - n0, ea[e0:e1]
        xc = x[n0:n1].double().requires_grad_(dX is not None)
        emb = torch.cat([t0.double()[a[:, :1]], tk.double()[a[:, 1:]]], dim=1)
        msg = (xc[src] + emb).masked_fill(a.unsqueeze(-1) == 0, 0.0)
        agg = torch.zeros_like(xc).index_add_(0, dst, msg)
        ref = ((torch.nn.functional.gelu(agg) + P[n0:n1].double()) * th.double()).sum(1)
        err = float((out[n0:n1].double() - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        assert err < 1e-5, "aggregation launch fails its check before timing: forward rel err %.3e" % err
        if dX is not None:
            ref.backward(dout[n0:n1].double())
            errb = float((dX[n0:n1].double() - xc.grad).abs().max() / xc.grad.abs().max())
            worst = max(worst, errb)
            assert errb < 1e-5, "aggregation launch fails its check before timing: dX rel err %.3e" % errb
    return worst


def agg_roofline(device, num_graphs, peak, reps=10):
    """The dominant kernel (fused K-hop aggregation forward, k=8, d=104, GELU + P + geometric combine) timed
    alone with CUDA events; algorithmic bytes per SURVEY.md 8(d):
    4*N*k*d [X] + 4*N*k*d [P] + 4*N*d [out] + 4*(N*k+1) [rowptr] + nnz*(4 [col] + 2 [attr16])."""
    import ctypes as C
    from kpgnn_b200 import _lib
    from kpgnn_b200.ops import _make_desc, want_blocks, ACT_GELU
    from kpgnn_b200.plan import get_plan
    hb = host_batch(num_graphs, seed=1000 + num_graphs)
    ei, ea = hb.edge_index.to(device), hb.edge_attr.to(device)
    N = hb.x.size(0)
    plan, k = get_plan(ei, ea, N)
    want_blocks(plan, k, HIDDEN, True)       # what khop_aggregate does: block-resident backward for large batches
    g = torch.Generator(device=device).manual_seed(0)
    x = torch.randn(N, K, HIDDEN, device=device, generator=g)
    P = torch.randn(N, K, HIDDEN, device=device, generator=g)
    t0 = torch.randn(5, HIDDEN, device=device, generator=g)
    tk = torch.randn(52, HIDDEN, device=device, generator=g)
    th = torch.softmax(torch.randn(K, HIDDEN, device=device, generator=g), 0)
    out = torch.empty(N, HIDDEN, device=device)
    desc = _make_desc(plan, k, x, P, t0, tk, th, None, ACT_GELU, True, False, False)
    lib = _lib.lib()
    st = torch.cuda.current_stream(device)
    sp = C.c_void_p(st.cuda_stream)
    alg = 4 * N * K * HIDDEN * 2 + 4 * N * HIDDEN + 4 * (N * K + 1) + plan.nnz * 6
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)     # 256 MB > 126 MB L2
    # verify, then time: the SAME descriptor / launch geometry
    _lib.check(lib.kp_agg_forward(C.byref(desc), out.data_ptr(), sp), "kp_agg_forward")
    checked = _check_agg_launch(hb, ei, ea, x, P, t0, tk, th, out)
    ts = []
    for i in range(reps + 3):
        flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        _lib.check(lib.kp_agg_forward(C.byref(desc), out.data_ptr(), sp), "kp_agg_forward")
        b.record(st)
        torch.cuda.synchronize(device)
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ms = statistics.mean(ts)
    achieved = alg / (ms * 1e-3) / 1e9
    # backward of the same call (B1 recompute + B2 transposed gather + B3 table gradients + reductions), timed as
    # one unit: algorithmic bytes per SURVEY 8(d): dOut + X (recompute) + dX + dP + P (dtheta) + index arrays twice
    dout = torch.randn(N, HIDDEN, device=device, generator=g)
    dX = torch.empty(N, K, HIDDEN, device=device)
    dP = torch.empty(N, K, HIDDEN, device=device)
    dT0, dTk, dth = torch.empty_like(t0), torch.empty_like(tk), torch.empty_like(th)
    from kpgnn_b200 import ops as OPS
    alg_b = 4 * N * HIDDEN + 4 * N * K * HIDDEN * 4 + 2 * (4 * (N * K + 1) + plan.nnz * 6)
    chunks = OPS.backward_chunks(plan, K, HIDDEN)       # node ranges of whole graphs when Gs [N,k,d] exceeds the L2

    def run_bwd(ch):
        OPS.agg_backward(plan, desc, dout, dX, dP, dT0, dTk, dth, None, chunks=ch)

    def time_bwd(ch):
        run_bwd(ch)
        err = _check_agg_launch(hb, ei, ea, x, P, t0, tk, th, out, dout, dX)
        ts = []
        if ch is not None and os.environ.get("KP_BENCH_GRAPH_BWD") == "1":      # experiment: device time without host launch cost
            run_bwd(ch)
            torch.cuda.synchronize(device)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                run_bwd(ch)
            for i in range(reps + 3):
                flush_l2(flush)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(torch.cuda.current_stream(device))
                gr.replay()
                b.record(torch.cuda.current_stream(device))
                torch.cuda.synchronize(device)
                if i >= 3:
                    ts.append(a.elapsed_time(b))
            return statistics.mean(ts), err
        for i in range(reps + 3):
            flush_l2(flush)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            run_bwd(ch)
            b.record(st)
            torch.cuda.synchronize(device)
            if i >= 3:
                ts.append(a.elapsed_time(b))
        return statistics.mean(ts), err
    ms_whole, err_w = time_bwd(None)
    checked = max(checked, err_w)
    tb = [ms_whole]
    ms_chunked = None
    if chunks is not None:
        ref_tabs = (dT0.clone(), dTk.clone(), dth.clone())
        ms_chunked, err_c = time_bwd(chunks)
        checked = max(checked, err_c)
        for got, want in zip((dT0, dTk, dth), ref_tabs):          # table / theta gradients: chunk sums vs the single call
            e = float((got - want).abs().max() / want.abs().max())
            assert e < 1e-5, "chunked backward: table / theta gradient differs from the single call by %.3e" % e
            checked = max(checked, e)
        tb = [ms_chunked]
    msb = statistics.mean(tb)
    del flush
    traffic = None
    tp = os.path.join(ROOT, "profiles", "agg_fwd_traffic.json")
    if os.path.isfile(tp):
        t = json.load(open(tp))
        if t.get("graphs_per_launch") == num_graphs:
            traffic = t["dram_bytes_per_launch"]
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "kernel": "agg_fwd_lean_kernel<32,GELU,fuse,smem-tables>",
            "graphs_per_launch": num_graphs, "nodes": N, "nnz": plan.nnz, "algorithmic_bytes": alg,
            "ms_per_launch": round(ms, 5), "checked_before_timing_max_rel_err": float("%.3e" % checked),
            "backward": {"ms": round(msb, 5), "algorithmic_bytes": alg_b,
                         "achieved": round(alg_b / (msb * 1e-3) / 1e9, 1),
                         "frac": round(alg_b / (msb * 1e-3) / 1e9 / peak, 4),
                         "kernels": ("agg_block_bwd_kernel (recompute + dP + dtheta, dX and table gradients from the "
                                     "shared-memory hand-over tile) + partial reductions") if plan.block_ptr is not None
                         else "agg_bwd_dst_lean (B1) + agg_fwd_lean<gather> (B2) + agg_bwd_table_count (B3) + reductions",
                         "node_range_chunks": None if chunks is None else len(chunks) - 1,
                         "ms_single_call": round(ms_whole, 5),
                         "ms_chunked": None if ms_chunked is None else round(ms_chunked, 5)}}


# ----------------------------------------------------------------------------------------------------------------
# other BASELINE.json workloads, reported next to the headline line (keys under "workloads")
# ----------------------------------------------------------------------------------------------------------------
REG_N, REG_K, REG_D, REG_GRAPHS = 1280, 6, 16, 64             # configs[4]: run_simulation.py:96-140 (n, K, hidden)
REG_EXTRACT = (REG_K, 10, 1, 1, 1, 1, "spd")                  # run_simulation.py:103


def _events_ms(fn, reps, device, flush=None, warm=2):
    st = torch.cuda.current_stream(device)
    ts = []
    for i in range(reps + warm):
        if flush is not None:
            flush_l2(flush)
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        torch.cuda.synchronize(device)
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return statistics.mean(ts)


def extract_alg_bytes(csr, out, args):
    """input int32 CSR + type (2*E1*4 + E1), output compact src/dst (E_K*8) + hop/attr per entry (nnz*3)
    + the peripheral tensors as int16 (N*K*(2*max_edge_type + max_hop_num + 1)*2)."""
    K, _, H, MET = args[0], args[1], args[2], args[3]
    E1 = int(csr["ecol"].shape[0])
    EK = int(out["edge_index"].size(1))
    nnz = int((out["edge_attr"] != 0).sum())
    return 2 * E1 * 4 + E1 + EK * 8 + nnz * 3 + csr["N"] * K * (2 * MET + H + 1) * 2


def extraction_metrics(device, peak, name, graphs, args, reps=5):
    """K-hop + peripheral extraction of one batch from its packed CSR (resident on the device): graphs/s, algorithmic
    GB/s and the fraction of the HBM roofline; bit-exactness is the gate (tests/test_extract_gpu.py)."""
    from kpgnn_b200 import data_utils as DU
    t0 = time.perf_counter()
    csr = DU.upload_csr(DU.pack_csr(graphs), device)
    pack_ms = (time.perf_counter() - t0) * 1e3
    holder = {}

    def run():
        holder["out"] = DU._extract_device(csr, *args, device=device)
    ms = _events_ms(run, reps, device)
    alg = extract_alg_bytes(csr, holder["out"], args)
    # the user-facing call on RAW host graphs: host CSR pack + one staged upload + kernels + output allocation, wall clock
    for _ in range(2):
        DU.extract_batch(graphs, args, device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(reps):
        DU.extract_batch(graphs, args, device)
    torch.cuda.synchronize(device)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
    return {"config": name, "ms_e2e_from_raw_graphs": round(e2e_ms, 3),
            "graphs_per_s_e2e": round(len(graphs) / (e2e_ms * 1e-3), 1), "graphs": len(graphs), "nodes": csr["N"], "khop_edges": int(holder["out"]["edge_index"].size(1)),
            "ms_device": round(ms, 3), "graphs_per_s": round(len(graphs) / (ms * 1e-3), 1),
            "algorithmic_bytes": alg, "achieved_GBps": round(alg / ms / 1e6, 2),
            "frac_of_hbm_peak": round(alg / ms / 1e6 / peak, 5), "host_csr_pack_ms": round(pack_ms, 3)}


class RegularWorkload(object):
    """configs[4]: node-level KP-GIN (`KGINConv`, run_simulation.py:29-93) on 3-regular graphs with n = 1 280, K = 6,
    hidden 16, forward only -- per step: K-hop extraction of the rank's 64 graphs (kp_extract_*), plan build, one
    KGINConv forward.  The reference runs the same per graph with batch_size 1 (run_simulation.py:103-110)."""

    def __init__(self, device, rank):
        from kpgnn_b200 import data_utils as DU, synth
        from kpgnn_b200.simulation import KGINConv
        self.DU, self.device = DU, device
        graphs = [synth.regular_graph(REG_N, 3, rank * REG_GRAPHS + s) for s in range(REG_GRAPHS)]
        self.csr = DU.pack_csr(graphs)
        self.dev_csr = DU.upload_csr(dict(self.csr), device)      # "value": the packed raw batch is resident in HBM
        torch.manual_seed(0)
        self.model = KGINConv(REG_D, REG_K).to(device).eval()
        self.x = torch.ones(self.csr["N"], 1, device=device)
        self.batch = torch.from_numpy(self.csr["node_graph"].astype(np.int64)).to(device)
        self.out = None
        self.host_out = torch.empty((self.csr["N"], REG_D), dtype=torch.float32, pin_memory=True)

    def step(self, csr=None):
        ex = self.DU._extract_device(csr or self.dev_csr, *REG_EXTRACT, device=self.device)
        with torch.no_grad():
            self.out = self.model(self.x, ex["edge_index"], ex["edge_attr"], self.batch)
        self.last = ex

    def step_e2e(self):
        self.step(self.csr)                                      # host arrays: uploaded inside the timed region
        self.host_out.copy_(self.out, non_blocking=True)        # run_simulation.py:111 `output.cpu()`
        torch.cuda.current_stream(self.device).synchronize()


def regular_roofline(device, peak, wl):
    """The aggregation launch of the configs[4] layer at >= 1 GB of algorithmic bytes: the rank's 64 extracted graphs
    replicated 10x with node offsets (640 graphs, 819 200 nodes, 144 M entries), d = 16, self term, no tables.
    Bytes per SURVEY.md 8(d): X + out + rowptr + 4-byte col per entry."""
    import ctypes as C
    from kpgnn_b200 import _lib
    from kpgnn_b200.ops import _make_desc, ACT_NONE
    from kpgnn_b200.plan import get_plan
    reps = 10
    ex = wl.last
    N0 = wl.csr["N"]
    ei = torch.cat([ex["edge_index"] + r * N0 for r in range(reps)], dim=1)
    ea = ex["edge_attr"].repeat(reps, 1)
    N = N0 * reps
    plan, k = get_plan(ei, ea, N)
    plan.blocks()
    del ei, ea
    x = torch.randn(N, REG_K, REG_D, device=device)
    eps = torch.zeros(1, device=device)
    out = torch.empty(N, REG_K, REG_D, device=device)
    desc = _make_desc(plan, k, x, None, None, None, None, eps, ACT_NONE, False, False, False)
    lib = _lib.lib()
    sp = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)
    ms = _events_ms(lambda: _lib.check(lib.kp_agg_forward(C.byref(desc), out.data_ptr(), sp), "kp_agg_forward"), 8, device,
                    flush)
    alg = 4 * N * REG_K * REG_D * 2 + 4 * (N * REG_K + 1) + plan.nnz * 4
    return {"bound": "hbm (algorithmic); the gathers themselves are served from shared memory", "graphs_per_launch":
            REG_GRAPHS * reps, "nodes": N, "nnz": plan.nnz, "algorithmic_bytes": alg, "ms_per_launch": round(ms, 4),
            "achieved": round(alg / ms / 1e6, 1), "peak": peak, "unit": "GB/s", "frac": round(alg / ms / 1e6 / peak, 4),
            "gather_model_GBps": round((plan.nnz * (REG_D * 4 + 4) + 4 * N * REG_K * REG_D) / ms / 1e6, 1),
            "kernel": "agg_tile_kernel<4,NONE,no-tables,self-term> (block-resident, csrc/agg_tile.cu)"}


def other_model_steps(device):
    """Training-step times (eager, fwd + bwd + Adam) of the other model configs on the product backbones, and -- when the
    reference's files are staged (oracle/_ref) -- of the reference's UNMODIFIED models/GNNs.py running layer by layer on
    the drop-in layers, next to the headline's stack path.  Diagnostics, not bench lines."""
    from kpgnn_b200 import backbones, synth
    from kpgnn_b200.data_utils import extract_batch
    from kpgnn_b200.model import l1_loss
    from kpgnn_b200.optim import FusedAdam
    res = {}
    graphs = synth.zinc_like_graphs(GRAPHS_PER_GPU, seed=0)
    y = torch.tensor([g["y"] for g in graphs], dtype=torch.float32, device=device)

    def time_model(model, b, steps=10):
        opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=1e-3)

        def step():
            opt.zero_grad()
            loss = l1_loss(model(b), y)
            loss.backward()
            opt.step()
        ms_eager = _events_ms(step, steps, device, warm=3)
        ms_graph = None
        try:                                    # the same step as ONE CUDA graph (kpgnn_b200.train.GraphedStep)
            from kpgnn_b200.train import GraphedStep
            gs = GraphedStep(model, b, y, l1_loss)
            ms_graph = _events_ms(gs.replay, steps, device, warm=3)
        except Exception as e:                  # diagnostics only
            log("GraphedStep failed: %r" % (e,))
        return ms_eager, ms_graph
    # configs[2]: KPGINPrime K=16, 17 layers, hidden 96 (README.md:128)
    b16 = extract_batch(graphs, (16, 50, 6, 3, 50, 50, "spd"), device)
    torch.manual_seed(0)
    prime = backbones.make_model("KPGINPrime", 96, 16, 17, 21, 3, 50, 50, 6, 50, JK="concat", residual=True).to(device).train()
    ms, msg = time_model(prime, b16)
    res["kpginprime_K16_L17_H96_batch128"] = {"ms_per_step_eager": round(ms, 3),
                                              "ms_per_step_cuda_graph": None if msg is None else round(msg, 3),
                                              "graphs_per_s": round(GRAPHS_PER_GPU / ((msg or ms) * 1e-3), 1)}
    try:
        from oracle import refimport
        if refimport.available():
            import argparse
            from kpgnn_b200.layers import layer_utils
            from kpgnn_b200.layers.input_encoder import EmbeddingEncoder
            dropin = refimport.load_models_over_dropin()
            b8 = extract_batch(graphs, EXTRACT_ARGS, device)
            a = argparse.Namespace(model_name="KPGINPlus", hidden_size=HIDDEN, K=K, num_hop1_edge=3, max_pe_num=50,
                                   combine="geometric", num_layer=LAYERS, eps=0., train_eps=False, aggr="add")
            torch.manual_seed(0)
            gnn = dropin.GNNs.GNNPlus(num_layer=LAYERS, gnn_layer=layer_utils.make_gnn_layer(a), JK="concat",
                                      norm_type="Batch", init_emb=EmbeddingEncoder(21, HIDDEN), residual=True,
                                      virtual_node=False, use_rd=False, num_hop1_edge=3, max_edge_count=50,
                                      max_hop_num=6, max_distance_count=50, wo_peripheral_edge=False,
                                      wo_peripheral_configuration=False, drop_prob=0.0)
            model = dropin.GraphRegression.GraphRegression(embedding_model=gnn, pooling_method="sum").to(device).train()
            ms, msg = time_model(model, b8)
            res["reference_GNNPlus_unmodified_over_dropin_layers_batch128"] = {
                "ms_per_step_eager": round(ms, 3), "ms_per_step_cuda_graph": None if msg is None else round(msg, 3),
                "graphs_per_s": round(GRAPHS_PER_GPU / ((msg or ms) * 1e-3), 1),
                "note": "the reference's own models/GNNs.py + GraphRegression.py (staged copy), layer by layer, no stack "
                        "node: what a reference user gets from install_dropin() alone (eager), and with the step "
                        "captured by kpgnn_b200.train.GraphedStep"}
    except Exception as e:      # diagnostics must never take the bench line down
        res["reference_GNNPlus_unmodified_over_dropin_layers_batch128"] = {"error": repr(e)[:200]}
    return res


def oracle_batch(num_graphs, seed):
    """The same synthetic batch as host_batch(), built WITHOUT the product: oracle extraction (numpy restatement of
    data_utils.py:20-241) per graph, then PyG Batch.from_data_list collation.  Untimed setup of the reference arm."""
    import numpy as np
    from kpgnn_b200 import synth
    from oracle.extract_np import extract_multi_hop_neighbors_np
    graphs = synth.zinc_like_graphs(num_graphs, seed=seed)
    keys = ("edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr")
    cols = {k: [] for k in keys}
    ei, xs, batch, off = [], [], [], 0
    for i, g in enumerate(graphs):
        o = extract_multi_hop_neighbors_np(g["num_nodes"], g["edge_index"], g["edge_attr"], *EXTRACT_ARGS)
        ei.append(o["edge_index"] + off)
        for k in keys:
            cols[k].append(o[k])
        xs.append(g["x"])
        batch.append(np.full(g["num_nodes"], i, dtype=np.int64))
        off += g["num_nodes"]
    b = {k: torch.from_numpy(np.concatenate(v, 0)) for k, v in cols.items()}
    b["edge_index"] = torch.from_numpy(np.concatenate(ei, 1))
    b["x"] = torch.from_numpy(np.concatenate(xs))
    b["batch"] = torch.from_numpy(np.concatenate(batch))
    b["y"] = torch.tensor([g["y"] for g in graphs], dtype=torch.float32)
    b["num_graphs"] = num_graphs
    return b


def cpu_baseline(steps=4, warmup=1):
    """The reference training step on the host's CPU cores, all threads.  With the reference's files at hand
    (/root/reference, or the staged copy oracle/_ref/ made by oracle/fetch_ref.py) it is the UNMODIFIED reference run
    through the torch_geometric stand-in (kind "reference", oracle/ref_step.py: train_ZINC.py:29-83); otherwise the
    oracle port of the same step (kind "port", oracle/model_torch.py).  Returns (per-step seconds, kind, threads)."""
    from oracle import ref_step
    if ref_step.available():
        from kpgnn_b200 import synth
        step, info = ref_step.zinc_reference_trainer(synth.zinc_like_graphs(GRAPHS_PER_GPU, seed=0), K, LAYERS, HIDDEN)
        ts = []
        for i in range(warmup + steps):
            t = time.perf_counter()
            step()
            if i >= warmup:
                ts.append(time.perf_counter() - t)
        return ts, "reference", info["threads"]
    from oracle.model_torch import l1_loss, zinc_oracle_model
    torch.set_num_threads(os.cpu_count())
    b = oracle_batch(GRAPHS_PER_GPU, seed=0)
    torch.manual_seed(0)
    model = zinc_oracle_model(K, LAYERS, HIDDEN).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    ts = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        opt.zero_grad()
        loss = l1_loss(model(b), b["y"])
        loss.backward()
        opt.step()
        loss.item()
        if i >= warmup:
            ts.append(time.perf_counter() - t)
    return ts, "port", torch.get_num_threads()


CPU_WHAT = {"reference": "the UNMODIFIED reference (models/GNNs.py GNNPlus + layers/KPGINplus.py through the "
                         "torch_geometric stand-in; extraction by the reference's data_utils.py, untimed)",
            "port": "oracle port of the reference step (oracle/model_torch.py)"}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the step, every host thread, honouring --steps /
    --warmup (one step is ~0.6 s on 16 cores, so the default 30 + 5 run ends within half a minute)."""
    if rank != 0:
        return
    ts, kind, threads = cpu_baseline(steps=args.steps, warmup=args.warmup)
    ms = 1e3 * statistics.mean(ts)
    val = GRAPHS_PER_GPU / (ms * 1e-3)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": "graphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": round(val, 2), "unit": "graphs/s", "cores": cores, "kind": kind,
                         "sample": "%d optimisation steps (after %d warm-up) of one 128-graph batch: %s, torch CPU, "
                                   "%d threads" % (args.steps, args.warmup, CPU_WHAT[kind], threads)},
        "e2e": {"value": round(val, 2), "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, exchange=None):
    cfg = {"workload": "configs[1]: ZINC-shape synthetic molecules, KPGINPlus K=8 8 layers hidden 104 residual, "
                        "batch 128 per GPU, spd kernel; forward+backward+Adam(lr 1e-3), L1 loss",
            "graphs_per_gpu": GRAPHS_PER_GPU, "global_batch": GRAPHS_PER_GPU * world,
            "parallelism": "dp%d" % world, "l2": "flushed between timed steps (256 MB write)",
            "plan_rebuilt_every_step": True, "cuda_graph": True,
            "batches": "%d distinct seeded batches per GPU (different node / K-hop edge / plan-entry counts), a "
                       "different one every step, through ONE captured step graph at padded capacity (row count "
                       "read from device memory)" % NUM_BATCHES,
            "e2e_input_pipeline": "every step uploads the next batch in the compact wire format (int32 ids, 1-byte "
                                  "attributes; ~1.4 MB instead of the 7.5 MB int64 layout) from pinned host memory on a "
                                  "copy stream while the previous step computes; device-to-device hand-over, one "
                                  "kernel widens it to the reference's int64 wire tensors; the loss of every step is copied to "
                                  "pinned host memory behind the step and read by the host one step late (while the next "
                                  "step runs), the deferred plan checks read sticky device-side maxima the same way"}
    if exchange:
        cfg["gradient_exchange"] = exchange
    return cfg


def main():
    import faulthandler
    faulthandler.dump_traceback_later(900, exit=True)      # a hang dumps every thread's stack and exits
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true",
                    help="skip the secondary workloads (configs[4] regular graphs, extraction metrics, other models)")
    ap.add_argument("--blas", default="default", choices=["default", "cublas", "cublaslt"])
    ap.add_argument("--eager", action="store_true",
                    help="profiling helper: no CUDA graph (ncu cannot re-launch graph kernel nodes that opted "
                         "into > 48 KB of dynamic shared memory); numbers from this mode are not bench values")
    ap.add_argument("--roofline-only", type=int, default=0, metavar="GRAPHS",
                    help="profiling helper: only time the aggregation kernel on GRAPHS graphs and exit")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    from kpgnn_b200 import build
    build.build_library()
    if os.environ.get("KPGNN_PROFILE_SMEM_CAP"):          # ncu launch lists only (see include/kpgnn.h)
        from kpgnn_b200 import _lib
        _lib.lib().kp_table_sum_set_smem_cap(int(os.environ["KPGNN_PROFILE_SMEM_CAP"]))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if args.blas != "default":
        torch.backends.cuda.preferred_blas_library(args.blas)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    peak, peak_src = peak_hbm()
    if args.roofline_only:
        print(json.dumps(agg_roofline(device, args.roofline_only, peak, reps=3)), flush=True)
        return

    log("[rank %d] building batch" % rank)
    bs = BenchStream(device, rank, world, eager=args.eager)
    tr = bs.tr
    log("[rank %d] captured, %d of our kernels per step; batches (N, E): %s" % (rank, tr.launches_per_step, bs.sizes))
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=device)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~0.3 s to start: begin before the warm-up steps
    for _ in range(args.warmup):
        bs.step_resident()
    torch.cuda.synchronize(device)
    tr.validate()
    t_res = timed_steps(bs.step_resident, args.steps, device, flush, dist_on)
    tr.validate()
    log("[rank %d] resident timing done" % rank)
    bs.tr.prefetch(bs.flats[(bs.i + 1) % NUM_BATCHES])
    for _ in range(3):
        bs.step_e2e()
    t_e2e = timed_steps(bs.step_e2e, args.steps, device, flush, dist_on)
    last_loss = tr.drain()                    # the last step's loss + the full deferred validation
    clocks = sampler.stop() if rank == 0 else None
    log("[rank %d] e2e timing done" % rank)

    def reduce_max(ms):
        if not dist_on:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())
    ms_res = reduce_max(statistics.mean(t_res))
    ms_e2e = reduce_max(statistics.mean(t_e2e))
    total_graphs = GRAPHS_PER_GPU * world

    # ---- configs[4]: regular n = 1280 graphs, extraction + KGIN forward, 64 graphs per GPU, at every N
    workloads = {}
    if not args.no_workloads:
        log("[rank %d] regular-graph workload" % rank)
        wl = RegularWorkload(device, rank)
        for _ in range(2):
            wl.step()
        t_reg = timed_steps(wl.step, max(3, min(args.steps, 10)), device, flush, dist_on)
        t_reg_e2e = timed_steps(wl.step_e2e, max(3, min(args.steps, 10)), device, flush, dist_on)
        ms_reg, ms_reg_e2e = reduce_max(statistics.mean(t_reg)), reduce_max(statistics.mean(t_reg_e2e))
        if rank == 0:
            nbytes_in = sum(int(np.asarray(wl.csr[k]).nbytes) for k in ("gptr", "node_graph", "pair_off", "erow", "ecol",
                                                                        "emult", "etype"))
            workloads["regular1280"] = {
                "workload": "configs[4]: 3-regular graphs n=1280, K=6 spd extraction + plan + KGINConv(16) forward, "
                            "%d graphs per GPU (run_simulation.py:96-116)" % REG_GRAPHS,
                "value": round(REG_GRAPHS * world / (ms_reg * 1e-3), 1), "unit": "graphs/s", "n_gpus": world,
                "ms_per_step": round(ms_reg, 3), "scaling": "weak",
                "e2e": {"value": round(REG_GRAPHS * world / (ms_reg_e2e * 1e-3), 1), "unit": "graphs/s",
                        "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": int(wl.host_out.numel() * 4),
                        "ms_per_step": round(ms_reg_e2e, 3)}}
            if world == 1 and not args.no_roofline:
                workloads["regular1280"]["roofline"] = regular_roofline(device, peak, wl)
        del wl
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_workloads:
        from kpgnn_b200 import synth
        log("[rank 0] extraction metrics / other model steps")
        zg = synth.zinc_like_graphs(GRAPHS_PER_GPU, seed=0)
        workloads["extract"] = [
            extraction_metrics(device, peak, "configs[1] zinc128 K=8 spd", zg, EXTRACT_ARGS),
            extraction_metrics(device, peak, "configs[2] zinc128 K=16 spd", zg, (16, 50, 6, 3, 50, 50, "spd")),
            # the same extraction at the roofline launch's size: what the kernels reach when they are not latency-bound
            extraction_metrics(device, peak, "configs[1] zinc%d K=8 spd" % ROOFLINE_GRAPHS,
                               synth.zinc_like_graphs(ROOFLINE_GRAPHS, seed=1000 + ROOFLINE_GRAPHS), EXTRACT_ARGS, reps=3),
            extraction_metrics(device, peak, "configs[4] regular1280 x16 K=6 spd",
                               [synth.regular_graph(REG_N, 3, s) for s in range(16)], REG_EXTRACT, reps=3)]
        workloads["model_steps"] = other_model_steps(device)
        torch.cuda.empty_cache()

    roof = roof_small = None
    cpu = None
    if rank == 0 and world == 1 and not args.no_roofline:
        roof_small = agg_roofline(device, GRAPHS_PER_GPU, peak)
        roof = agg_roofline(device, ROOFLINE_GRAPHS, peak)
        roof["peak_source"] = peak_src
        roof["note"] = ("fused K-hop aggregation forward timed alone (CUDA events, L2 flushed) at a launch with "
                        ">= 1 GB algorithmic bytes as SURVEY.md 8(d) requires; roofline_batch128 is the same kernel "
                        "at the bench batch (21 MB, L2-resident / launch-bound)")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ts, kind, threads = cpu_baseline(steps=10, warmup=2)
        v = GRAPHS_PER_GPU / statistics.mean(ts)
        cpu = {"value": round(v, 2), "unit": "graphs/s", "cores": os.cpu_count(), "kind": kind,
               "sample": "10 optimisation steps (after 2 warm-up) of one 128-graph batch: %s, torch CPU, %d threads"
                         % (CPU_WHAT[kind], threads)}
    if dist_on:
        torch.distributed.barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(total_graphs / (ms_res * 1e-3), 1), "unit": "graphs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, None if world == 1 else (
                "peer-memory kernel over NVLink (csrc/peer.cu), captured in the step graph"
                if type(tr.grads).__name__ == "PeerGradients" else "process-group all-reduce between two graphs")),
            "e2e": {"value": round(total_graphs / (ms_e2e * 1e-3), 1), "unit": "graphs/s",
                    "h2d_bytes_per_step": int(bs.spec.nbytes), "d2h_bytes_per_step": 4 + 16,
                    "ms_per_step": round(ms_e2e, 4),
                    "reference_wire_layout_bytes_per_batch": bs.wire_bytes_reference_layout},
            "gpu_launches": int(tr.launches_per_step * args.steps),
            "gpu_launches_per_step": {"ours_in_cuda_graph_incl_plan_rebuild": int(tr.launches_per_step)},
            "clocks": clocks, "roofline": roof, "roofline_batch128": roof_small, "cpu_baseline": cpu,
            "workloads": workloads or None, "loss": float(tr.loss.item()),
        }
        print(json.dumps(line), flush=True)
    if dist_on:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
