"""Kineto timeline of ONE replay of the captured training step (bench.py's BenchStream): start (us from the step's first
kernel), duration, stream, kernel name.  `python profiles/step_timeline.py > gpurun_out/step_timeline.txt`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    bs = bench.BenchStream(dev, 0, 1)
    for _ in range(5):
        bs.step_resident()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            bs.step_resident()
        torch.cuda.synchronize()
    # raw Kineto records carry the stream id (device_resource_id); FunctionEvent does not
    evs = [e for e in prof.profiler.kineto_results.events()
           if e.device_type() == torch.autograd.DeviceType.CUDA and e.duration_ns() > 0]
    evs.sort(key=lambda e: e.start_ns())
    starts = [i for i, e in enumerate(evs) if "wire_unpack" in e.name()]
    seg = evs[starts[-1]:]
    t0 = seg[0].start_ns()
    end = max(e.start_ns() + e.duration_ns() for e in seg)
    print("# one step: %d device activities, %.1f us from first start to last end" % (len(seg), (end - t0) / 1e3))
    tot = {}
    for e in seg:
        dur = e.duration_ns() / 1e3
        print("%8.1f %7.1f  s%-3d %s" % ((e.start_ns() - t0) / 1e3, dur, e.device_resource_id(), e.name()[:110]))
        key = e.name().split("(")[0][:60]
        tot[key] = tot.get(key, 0.0) + dur
    print("# totals")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
        print("# %8.1f us  %s" % (v, k))


if __name__ == "__main__":
    main()
