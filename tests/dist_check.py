"""Run under torchrun on >= 2 GPUs (tests/test_dist_gpu.py launches it): the data-parallel step of kpgnn_b200/train.py
on hardware -- the all-reduced flat gradient equals the mean of the ranks' local gradients, and replicas that start equal
stay bit-identical over optimisation steps on different per-rank batches (captured step graph + NCCL all-reduce)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    logf = open(os.path.join(ROOT, "gpurun_out", "dist_rank%d.log" % rank), "a")

    def mark(msg):
        logf.write("[in_graph=%s] %s\n" % (os.environ.get("KP_NCCL_IN_GRAPH", "0"), msg))
        logf.flush()
    mark("start")
    same_gpu = os.environ.get("KP_DIST_SAME_GPU") == "1"      # two processes on ONE device (single-GPU boxes): gloo plumbing
    dev = torch.device("cuda", 0 if same_gpu else local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)

    def all_gather(t):
        if same_gpu:
            out = [torch.empty_like(t, device="cpu") for _ in range(world)]
            dist.all_gather(out, t.cpu())
            return [o.to(dev) for o in out]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return out
    mark("pg up")
    from kpgnn_b200 import synth
    from kpgnn_b200.data_utils import extract_batch_host
    from kpgnn_b200.model import Batch, zinc_kpginplus
    from kpgnn_b200.train import Trainer, fit_spec
    args = (8, 50, 6, 3, 50, 50, "spd")
    hbs = []
    for s in range(3):
        graphs = synth.zinc_like_graphs(32, seed=100 * rank + s)
        f = extract_batch_host(graphs, args, dev)
        f["y"] = torch.tensor([g["y"] for g in graphs], dtype=torch.float32)
        hbs.append(Batch(**f))
    spec, bounds = fit_spec(hbs, 8, 3, 6, headroom=1.3)
    flats = [spec.pack(b, spec.host_buffer()) for b in hbs]
    mark("batches packed")
    # (1) eager: all-reduced gradient == mean of the local gradients
    torch.manual_seed(0)
    model = zinc_kpginplus(8, 8, 104).to(dev).train()
    tr = Trainer(model, spec, bounds, dev, world=world, use_graph=False)
    tr.load(flats[0])
    tr._fwd_bwd()
    mark("eager fwd/bwd")
    peer = hasattr(tr.grads, "send")
    local_flat = (tr.grads.send if peer else tr.grads.flat).clone()
    tr.grads.allreduce_mean_(world)
    gathered = all_gather(local_flat)
    torch.cuda.synchronize()
    mark("all-reduce + all-gather")
    mean = torch.stack(gathered).double().mean(0)
    err = float((tr.grads.flat.double() - mean).abs().max() / mean.abs().max())
    if peer:        # the kernel's contract: fp32 sum in rank order, times 1/world -- bit-exact, identical on every rank
        acc = gathered[0].clone()
        for g in gathered[1:]:
            acc += g
        assert torch.equal(tr.grads.flat, acc * (1.0 / world)), "peer exchange is not the rank-ordered fp32 mean"
    differ = float((gathered[0] - gathered[-1]).abs().max())
    assert err < 1e-6, err
    assert differ > 0, "ranks must see different batches"
    # (2) captured step graph: replicas stay identical over steps on different per-rank batches
    torch.manual_seed(0)
    model = zinc_kpginplus(8, 8, 104).to(dev).train()
    tr = Trainer(model, spec, bounds, dev, world=world, use_graph=True)
    tr.capture(flats[0])
    mark("captured, single_graph=%s" % (tr.graph_opt is None))
    tr.prefetch(flats[0])
    for step in range(6):
        tr.step_e2e(flats[(step + 1) % 3])
        mark("step %d" % step)
    flat_params = torch.cat([p.detach().flatten() for p in model.parameters()])
    allp = all_gather(flat_params)
    if peer:
        tr.grads.check()
    drift = max(float((a - allp[0]).abs().max()) for a in allp)
    assert drift == 0.0, drift
    if rank == 0:
        print("DIST_CHECK_OK world=%d exchange=%s allreduce_err=%.2e replicas_drift=%.1f single_graph=%s"
              % (world, "peer-memory kernel" if peer else "process group", err, drift, tr.graph_opt is None))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
