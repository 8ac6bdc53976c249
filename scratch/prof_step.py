"""Kineto (torch.profiler) view of the bench training step: (1) CUDA-graph replay -> real in-step kernel durations
and idle gaps; (2) eager step -> which autograd node / aten op launches each glue kernel.  Not a bench value."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def short(n):
    n = n.replace("void ", "").replace("at::native::", "").replace("at::", "")
    return n[:110]


def graph_profile():
    tr = bench.Trainer(bench.host_batch(bench.GRAPHS_PER_GPU, seed=0), dev, 1)
    tr.capture()
    for _ in range(5):
        tr.step_resident()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            tr.step_resident()
            torch.cuda.synchronize()
    ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks.sort(key=lambda e: e.time_range.start)
    if not ks:
        print("no CUDA events")
        return
    # split into steps by large gaps
    steps, cur = [], [ks[0]]
    for a, b in zip(ks[:-1], ks[1:]):
        if b.time_range.start - a.time_range.end > 200:
            steps.append(cur)
            cur = []
        cur.append(b)
    steps.append(cur)
    st = steps[-1]
    span = st[-1].time_range.end - st[0].time_range.start
    busy = sum(e.time_range.end - e.time_range.start for e in st)
    print("GRAPH step: %d kernels, span %.1f us, busy %.1f us, idle %.1f us" % (len(st), span, busy, span - busy))
    agg = collections.OrderedDict()
    for e in st:
        k = short(e.name)
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += e.time_range.end - e.time_range.start
        agg[k][1] += 1
    for k, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:45]:
        print("%8.1f us %4d x %6.2f  %s" % (t, c, t / c, k))
    # timeline of the last replay: start offset, duration, stream, kernel
    per = len(st) // 3
    last = st[-per:]
    t0 = last[0].time_range.start
    print("\nTIMELINE of one step (%d kernels, %.1f us):" % (len(last), last[-1].time_range.end - t0))
    for e in last:
        print("%8.1f %6.1f  s%-3s %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start,
                                        getattr(e, "device_resource_id", "?"), short(e.name)[:70]))
    del tr


def eager_profile():
    tr = bench.Trainer(bench.host_batch(bench.GRAPHS_PER_GPU, seed=0), dev, 1, eager=True)
    tr.capture()
    for _ in range(3):
        tr.step_resident()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step_resident()
        torch.cuda.synchronize()
    evs = prof.events()
    cnt = collections.Counter()
    tot = collections.Counter()
    for e in evs:
        if e.device_type != torch.autograd.DeviceType.CPU:
            continue
        kern = [k for k in e.kernels] if hasattr(e, "kernels") else []
        if not kern:
            continue
        # only leaf ops that own kernels directly
        if any(c.kernels for c in e.cpu_children if hasattr(c, "kernels")):
            continue
        chain = []
        p = e
        while p is not None:
            chain.append(p.name)
            p = p.cpu_parent
        top = chain[-1]
        key = (e.name, top if top != e.name else (chain[1] if len(chain) > 1 else ""))
        cnt[key] += len(kern)
        tot[key] += sum(k.duration for k in kern)
    print("\nEAGER step: kernels by (leaf op, outermost parent)")
    for key, t in sorted(tot.items(), key=lambda x: -x[1])[:70]:
        print("%8.1f us %4d  %-28s <- %s" % (t, cnt[key], key[0][:28], key[1][:90]))


graph_profile()
if os.environ.get('KP_PROF_EAGER'):
    eager_profile()
