"""Where does the e2e - resident gap go?  Host-timer breakdown of step_e2e and variants (not a bench value)."""
import os, sys, time, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.backends.cuda.matmul.allow_tf32 = False
tr = bench.Trainer(bench.host_batch(bench.GRAPHS_PER_GPU, seed=0), dev, 1)
tr.capture()
for _ in range(5): tr.step_resident()
for _ in range(5): tr.step_e2e()
torch.cuda.synchronize()
def gpu_time(fn, n=30):
    ts = []
    st = torch.cuda.current_stream(dev)
    for _ in range(n):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.mean(ts) * 1e3
print("resident (replay only)            %.1f us" % gpu_time(tr.step_resident))
def v_item():
    tr.replay(); tr.loss.item()
print("replay + loss.item()              %.1f us" % gpu_time(v_item))
def v_handover():
    tr.hand_over(); tr.replay(); tr.loss.item()
tr.prefetch(); torch.cuda.synchronize()
print("hand_over + replay + item         %.1f us" % gpu_time(v_handover))
def v_full_noval():
    tr.hand_over(); tr.replay(); tr.prefetch(); tr.loss.item()
print("+ prefetch (H2D overlapped)       %.1f us" % gpu_time(v_full_noval))
print("step_e2e                          %.1f us" % gpu_time(lambda: tr.step_e2e()))
# host-side cost of each call
def host(fn, n=50):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
print("host: hand_over %.1f us, prefetch %.1f us, validate %.1f us" % (host(tr.hand_over), host(tr.prefetch), host(lambda: tr.plan().validate())))
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
cs = tr.copy_stream
torch.cuda.synchronize()
with torch.cuda.stream(cs):
    a.record(cs); tr.stage_flat.copy_(tr.host_flat, non_blocking=True); b.record(cs)
torch.cuda.synchronize()
print("H2D of %.1f MB alone: %.1f us" % (tr.host_flat.numel() / 1e6, a.elapsed_time(b) * 1e3))
# host-timer trace of the pieces inside one e2e step
import collections
acc = collections.defaultdict(list)
for it in range(40):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.hand_over()
    t1 = time.perf_counter(); tr.replay()
    t2 = time.perf_counter(); tr.prefetch()
    t3 = time.perf_counter(); v = tr.loss.item()
    t4 = time.perf_counter(); p = tr.plan_obj
    t5 = time.perf_counter(); p.validate()
    t6 = time.perf_counter()
    if it >= 10:
        for k, a_, b_ in (("hand_over", t0, t1), ("replay", t1, t2), ("prefetch", t2, t3), ("item", t3, t4), ("plan()", t4, t5), ("validate", t5, t6), ("total", t0, t6)):
            acc[k].append((b_ - a_) * 1e6)
print("host trace (us): " + ", ".join("%s %.1f" % (k, statistics.mean(v)) for k, v in acc.items()))
