"""GPU parity of the plan builder (integer work: bit-exact against a numpy construction)."""
import numpy as np
import pytest
import torch

from tests.util import zinc_batch

pytestmark = pytest.mark.gpu


def _expected(ei, ea, N, K, self_loops):
    src, dst = ei[0], ei[1]
    rows, rowsT = [[] for _ in range(N * K)], [[] for _ in range(N * K)]
    for e in range(ei.shape[1]):
        for h in range(K):
            if ea[e, h] != 0:
                rows[dst[e] * K + h].append((int(src[e]), int(ea[e, h])))
                rowsT[src[e] * K + h].append(int(dst[e]))
    if self_loops:
        for v in range(N):
            for h in range(K):
                rows[v * K + h].append((v, 1))
                rowsT[v * K + h].append(v)
    rowptr = np.cumsum([0] + [len(r) for r in rows])
    rowptrT = np.cumsum([0] + [len(r) for r in rowsT])
    col = np.array([c for r in rows for c, _ in r], dtype=np.int64)
    att = np.array([a for r in rows for _, a in r], dtype=np.int64)
    colT = np.array([c for r in rowsT for c in r], dtype=np.int64)
    return rowptr, col, att, rowptrT, colT


@pytest.mark.parametrize("kern", ["spd", "gd"])
@pytest.mark.parametrize("self_loops", [False, True])
def test_plan_matches_numpy(lib, kern, self_loops):
    from kpgnn_b200.plan import get_plan
    b = zinc_batch(5, 4, kern, seed=3)
    dev = torch.device("cuda:0")
    ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)
    N, K = b["num_nodes"], 4
    plan, k = get_plan(ei, ea, N, self_loops)
    assert k == K and plan.K == K
    rowptr, col, att, rowptrT, colT = _expected(b["edge_index"].numpy(), b["edge_attr"].numpy(), N, K, self_loops)
    nnz = int(rowptr[-1])
    assert plan.nnz == nnz
    assert np.array_equal(plan.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(plan.rowptrT.cpu().numpy(), rowptrT)
    assert np.array_equal(plan.col.cpu().numpy()[:nnz], col)
    assert np.array_equal(plan.attr16.cpu().numpy()[:nnz].astype(np.int64) & 0xFFFF, att)
    assert np.array_equal(plan.colT.cpu().numpy()[:nnz], colT)
    indeg = np.bincount(b["edge_index"][1].numpy(), minlength=N)
    assert np.array_equal(plan.indeg.cpu().numpy(), indeg)
    if self_loops:
        deg = np.diff(rowptr).astype(np.float32)
        assert np.allclose(plan.dinv.cpu().numpy(), 1 / np.sqrt(deg), rtol=1e-6)
    # column-sliced views resolve to the same cached plan and serve k < K
    plan2, k2 = get_plan(ei, ea[:, :2], N, self_loops)
    assert plan2 is plan and k2 == 2


def test_plan_rejects_bad_indices(lib):
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    ei = torch.tensor([[0, 1, 5], [1, 0, 0]], device=dev)
    ea = torch.ones(3, 2, dtype=torch.long, device=dev)
    with pytest.raises(IndexError):
        get_plan(ei, ea, 3)


def test_plan_empty(lib):
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    ei = torch.zeros(2, 0, dtype=torch.long, device=dev)
    ea = torch.zeros(0, 3, dtype=torch.long, device=dev)
    plan, k = get_plan(ei, ea, 4)
    assert plan.nnz == 0 and k == 3
    assert int(plan.rowptr.abs().sum()) == 0


def test_plans_of_discarded_batches_are_freed_without_the_cyclic_gc(lib):
    """The plan is cached on the edge_index tensor object; it must not reference that tensor back (a view's ._base is the
    tensor itself), or every batch of a per-step collation loop (train_ZINC.py:33: a new Batch every step) stays allocated
    until the cyclic garbage collector happens to run."""
    import gc
    from kpgnn_b200.plan import get_plan
    dev = torch.device("cuda:0")
    b = zinc_batch(24, 6, "spd", seed=5)
    N = b["num_nodes"]
    gc.collect()
    gc.disable()
    try:
        sizes = []
        for step in range(6):
            ei, ea = b["edge_index"].to(dev), b["edge_attr"].to(dev)          # fresh tensor objects, as a loader delivers
            plan, _ = get_plan(ei, ea, N)
            plan.blocks()
            del plan, ei, ea
            torch.cuda.synchronize()
            sizes.append(torch.cuda.memory_allocated())
    finally:
        gc.enable()
    assert sizes[-1] <= sizes[1], sizes
