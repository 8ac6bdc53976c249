"""A/B of the dense block's GEMM phases (tensor-core 3xTF32 vs fp32-FMA tiles) and of the rows-per-CTA choice, at the
bench batch (N = 2 986 rows, 104 channels): CUDA-event time of forward and backward, 200 launches each, plus the error of
both paths against a float64 evaluation.  `KP_DENSE_ROWS=<r> python profiles/dense_ab.py`"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from kpgnn_b200 import _lib
    from kpgnn_b200.layers.dense_block import fused_dense_block
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    N, C = 2986, 104
    torch.manual_seed(0)
    mods = [nn.Linear(C, C), nn.BatchNorm1d(C), nn.Linear(C, C), nn.BatchNorm1d(C), nn.BatchNorm1d(C)]
    mods = [m.to(dev).train() for m in mods]
    x = torch.randn(N, C, device=dev, requires_grad=True)
    r = torch.randn(N, C, device=dev)
    gy = torch.randn(N, C, device=dev)
    # float64 truth
    m64 = [nn.Linear(C, C), nn.BatchNorm1d(C), nn.Linear(C, C), nn.BatchNorm1d(C), nn.BatchNorm1d(C)]
    m64 = [m.to(dev).double().train() for m in m64]
    for a, b in zip(m64, mods):
        a.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in b.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = m64[4](torch.relu(m64[3](m64[2](torch.relu(m64[1](m64[0](x64))))))) + r.double()
    y64.backward(gy.double())
    lib = _lib.lib()
    for mode, name in ((1, "mma-3xtf32"), (0, "fp32-fma")):
        lib.kp_dense_block_set_mma(mode)
        for m in mods:
            m.zero_grad()
        x.grad = None
        y = fused_dense_block(x, *mods, r)
        y.backward(gy)
        ey = float((y.double() - y64).abs().max() / y64.abs().max())
        ex = float((x.grad.double() - x64.grad).abs().max() / x64.grad.abs().max())
        ew = float((mods[0].weight.grad.double() - m64[0].weight.grad).abs().max() / m64[0].weight.grad.abs().max())
        ts = {}
        for what in ("fwd", "bwd"):
            tt = []
            for i in range(220):
                x.grad = None
                if what == "fwd":
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    y = fused_dense_block(x, *mods, r)
                    b.record()
                else:
                    y = fused_dense_block(x, *mods, r)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    y.backward(gy)
                    b.record()
                torch.cuda.synchronize()
                if i >= 20:
                    tt.append(a.elapsed_time(b) * 1e3)
            tt.sort()
            ts[what] = tt[len(tt) // 2]
        print("rows=%s %-11s fwd %.1f us  bwd(+autograd glue) %.1f us   err vs fp64: y %.2e dx %.2e dW1 %.2e"
              % (os.environ.get("KP_DENSE_ROWS", "default"), name, ts["fwd"], ts["bwd"], ey, ex, ew))
    lib.kp_dense_block_set_mma(-1)


if __name__ == "__main__":
    main()
