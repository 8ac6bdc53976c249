"""Diagnostic for a dense-block parity failure: are the rows whose dX differs from torch's the rows that sit on a ReLU
kink (|pre-activation| at rounding level), i.e. a mask flip between two fp32 evaluations, or is it a kernel bug?"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_dense_gpu import _modules  # noqa: E402


def main():
    from kpgnn_b200.layers.dense_block import fused_dense_block
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    for (N, Cin, Cout) in ((5000, 100, 100), (5000, 104, 104), (3, 128, 128)):
        g = torch.Generator().manual_seed(N + Cin)
        x0 = (torch.randn(N, Cin, generator=g) * 2 + 0.5).to(dev)
        r0 = torch.randn(N, Cout, generator=g).to(dev)
        gy = torch.randn(N, Cout, generator=g).to(dev)
        lin1, bn1, lin2, bn2, bn3 = _modules(Cin, Cout, 1, dev)
        x = x0.clone().requires_grad_(True)
        a1 = bn1(lin1(x))
        a2 = bn2(lin2(torch.relu(a1)))
        y = bn3(torch.relu(a2)) + r0
        y.backward(gy)
        dx_ref = x.grad.clone()
        mods = _modules(Cin, Cout, 1, dev)
        x2 = x0.clone().requires_grad_(True)
        y2 = fused_dense_block(x2, *mods, r0)
        y2.backward(gy)
        diff = (x2.grad - dx_ref).abs()
        scale = float(dx_ref.abs().max())
        bad_rows = (diff.max(dim=1).values / scale > 2e-5).nonzero().flatten()
        print("N=%d C=%d: max rel diff %.3e, rows above 2e-5: %d of %d" % (N, Cin, float(diff.max()) / scale, len(bad_rows), N))
        for r in bad_rows[:6].tolist():
            print("   row %d: rel diff %.2e, min|BN1 out| %.2e, min|BN2 out| %.2e  (fp32 eps at that scale ~1e-7)"
                  % (r, float(diff[r].max()) / scale, float(a1[r].abs().min()), float(a2[r].abs().min())))
        med = float(diff.median()) / scale
        print("   median rel diff over all elements %.2e; global min|BN1 out| %.2e min|BN2 out| %.2e"
              % (med, float(a1.abs().min()), float(a2.abs().min())))


if __name__ == "__main__":
    main()
