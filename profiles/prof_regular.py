"""Where does a configs[4] step (extraction of 64 n=1280 graphs + plan + KGINConv forward) spend its time?
CUDA-event time per step over 8 steps, then a CPU+CUDA profile of 3 steps."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    wl = bench.RegularWorkload(dev, 0)
    for _ in range(2):
        wl.step()
    torch.cuda.synchronize()
    for i in range(8):
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        wl.step()
        b.record()
        torch.cuda.synchronize()
        print("step %d: events %.2f ms, wall %.2f ms, reserved %.1f GB allocated %.1f GB" % (
            i, a.elapsed_time(b), (time.perf_counter() - t0) * 1e3, torch.cuda.memory_reserved() / 2**30,
            torch.cuda.memory_allocated() / 2**30))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            wl.step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=10, max_name_column_width=60))
    print(torch.cuda.memory_stats().get("num_alloc_retries"), torch.cuda.memory_stats().get("num_device_alloc"),
          torch.cuda.memory_stats().get("num_device_free"))


if __name__ == "__main__":
    main()
