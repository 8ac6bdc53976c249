"""Multi-table gather-sum (kp_table_sum_*, csrc/tsum.cu + tsum_sorted.cu) against plain torch indexing: forward and
the deterministic table gradient, for the peripheral-encoder slot layout (models/GNNs.py:393-400) and the 1-slot
input embedding, with uniform and heavily skewed indices (index 0 dominating, as in real peripheral attributes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layout(sizes, slots):
    from kpgnn_b200.encoders import _ranges
    offs, o = [], 0
    for n, s in zip(sizes, slots):
        offs += [o] * s
        o += n
    rs, rr = _ranges(sizes, slots, 104)
    return offs, rs, rr, o


@pytest.mark.parametrize("R", [1, 31, 2952, 23824])
@pytest.mark.parametrize("layout", ["peripheral", "input", "wide"])
@pytest.mark.parametrize("skew", [False, True])
def test_table_sum_matches_torch(lib, R, layout, skew):
    from kpgnn_b200.encoders import _TableSum
    dev = torch.device("cuda:0")
    sizes, slots = {"peripheral": ([5, 51] + [51] * 7 + [1], [3, 3] + [1] * 7 + [1]),
                    "input": ([21], [1]),
                    "wide": ([200, 7, 256], [2, 5, 1])}[layout]
    d = 104
    offs, rs, rr, rows = _layout(sizes, slots)
    g = torch.Generator().manual_seed(R + len(sizes))
    cols = []
    for n, s in zip(sizes, slots):
        for _ in range(s):
            v = torch.randint(0, n, (R,), generator=g)
            if skew:
                v = torch.where(torch.rand(R, generator=g) < 0.85, torch.zeros_like(v), v)
            cols.append(v)
    idx = torch.stack(cols, 1).to(dev)
    table0 = torch.randn(rows, d, generator=g).to(dev)
    gy = torch.randn(R, d, generator=g).to(dev)
    res = []
    for fused in (True, False):
        table = table0.clone().requires_grad_(True)
        if fused:
            out = _TableSum.apply(table, idx, offs, rs, rr)
        else:
            out = sum(table[offs[s] + idx[:, s]] for s in range(idx.size(1)))
        out.backward(gy)
        res.append((out.detach(), table.grad))
    for a, b in zip(*res):
        scale = max(float(b.abs().max()), 1e-6)
        assert float((a - b).abs().max()) / scale < 1e-5
    table = table0.clone().requires_grad_(True)               # bit-reproducible
    _TableSum.apply(table, idx, offs, rs, rr).backward(gy)
    assert torch.equal(table.grad, res[0][1])
