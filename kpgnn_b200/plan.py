"""Graph plan: the device-resident (dst,hop)-CSR / (src,hop)-CSR every layer of a batch shares.

Host-side wrapper over kp_plan_count / kp_plan_fill (include/kpgnn.h).  The reference has no such object: PyG's
`propagate` re-derives gather/scatter indices from `edge_index [2,E]` / `edge_attr [E,K]` in every layer
(layers/KPGIN.py:100 etc.).  A plan is built once per batch (one host sync, to size the compact arrays) and is
cached ON the `edge_index` tensor object, which `models/GNNs.py` hands unchanged to every layer
(GNNs.py:190,429,655,679); the column-sliced views `edge_attr[:, :k]` (GNNs.py:429,679) all resolve to the plan
of their base tensor, which serves every k <= K.
"""
import ctypes as C

import torch

from . import _lib


class GraphPlan(object):
    __slots__ = ("N", "E", "K", "nnz", "self_loops", "rowptr", "col", "attr16", "rowptrT", "colT", "dinv", "indeg",
                 "max_attr0", "max_attrk", "device")

    def check_tables(self, rows0, rowsk, k):
        """nn.Embedding would raise IndexError on an out-of-range attr (KPGIN.py:90,95); so do we."""
        if self.max_attr0 >= rows0:
            raise IndexError("edge_attr[:,0] has value %d but hop1_edge_emb has %d rows" % (self.max_attr0, rows0))
        if k > 1 and self.max_attrk >= rowsk:
            raise IndexError("edge_attr[:,1:] has value %d but hopk_edge_emb has %d rows" % (self.max_attrk, rowsk))


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def build_plan(edge_index, edge_attr_base, attr_stride, K, num_nodes, self_loops=False):
    """edge_index [2,E] int64 cuda; edge_attr_base: int64 cuda tensor whose element (e,h) lives at
    data_ptr + 8*(e*attr_stride + h) for h < K."""
    lib = _lib.lib()
    if not edge_index.is_cuda:
        raise _lib.KpError("kpgnn_b200 runs on CUDA tensors only (no CPU fallback); got edge_index on %s"
                           % edge_index.device)
    assert edge_index.dtype == torch.int64 and edge_attr_base.dtype == torch.int64
    dev = edge_index.device
    E = edge_index.size(1)
    N = int(num_nodes)
    src = edge_index[0].contiguous()
    dst = edge_index[1].contiguous()
    rows = N * K
    pin = _lib.PlanInput(src.data_ptr(), dst.data_ptr(), edge_attr_base.data_ptr(), attr_stride, N, E, K,
                         1 if self_loops else 0)
    nbytes = C.c_size_t(0)
    _lib.check(lib.kp_plan_workspace_bytes(N, E, K, C.byref(nbytes)), "kp_plan_workspace_bytes")
    ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
    p = GraphPlan()
    p.N, p.E, p.K, p.self_loops, p.device = N, E, K, bool(self_loops), dev
    p.rowptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    p.rowptrT = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    p.indeg = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    stats = torch.empty(4, dtype=torch.int32, device=dev)
    st = _stream_ptr(dev)
    _lib.check(lib.kp_plan_count(C.byref(pin), p.rowptr.data_ptr(), p.rowptrT.data_ptr(), p.indeg.data_ptr(),
                                 stats.data_ptr(), ws.data_ptr(), ws.numel(), st), "kp_plan_count")
    nnz, p.max_attr0, p.max_attrk, bad = stats.tolist()          # the one host sync per batch
    if bad:
        raise IndexError("edge_index / edge_attr out of range in %d entries (node ids must be in [0,%d), "
                         "attrs in [0,65535])" % (bad, N))
    p.nnz = nnz
    p.col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    p.colT = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    p.attr16 = torch.empty(max(nnz, 1), dtype=torch.int16, device=dev)
    p.dinv = torch.empty(max(rows, 1), dtype=torch.float32, device=dev) if self_loops else None
    _lib.check(lib.kp_plan_fill(C.byref(pin), p.rowptr.data_ptr(), p.rowptrT.data_ptr(), p.col.data_ptr(),
                                p.attr16.data_ptr(), p.colT.data_ptr(),
                                p.dinv.data_ptr() if self_loops else None, ws.data_ptr(), ws.numel(), st),
               "kp_plan_fill")
    return p


def _attr_base(edge_attr):
    """Resolve a (possibly column-sliced) edge_attr view to (tensor covering the base rows, stride, K_full)."""
    if edge_attr.dim() == 1:
        edge_attr = edge_attr.view(-1, 1)
    E, k = edge_attr.shape
    base = edge_attr._base
    if (base is not None and base.dim() == 2 and base.is_contiguous() and base.size(0) == E
            and edge_attr.stride(1) == 1 and edge_attr.stride(0) == base.size(1)
            and edge_attr.data_ptr() == base.data_ptr()):
        return base, base.size(1), base.size(1)
    if not edge_attr.is_contiguous():
        edge_attr = edge_attr.contiguous()
    return edge_attr, k, k


def get_plan(edge_index, edge_attr, num_nodes, self_loops=False):
    """Cached plan lookup.  Returns (plan, k) where k = edge_attr.size(1) hops of the plan are in use."""
    k = edge_attr.size(1) if edge_attr.dim() == 2 else 1
    base, stride, K = _attr_base(edge_attr)
    key = (base.data_ptr(), base._version, stride, K, int(num_nodes), bool(self_loops), edge_index._version)
    cache = getattr(edge_index, "_kpgnn_plans", None)
    if cache is None:
        cache = {}
        try:
            edge_index._kpgnn_plans = cache
        except Exception:      # pragma: no cover - tensors always accept attributes
            pass
    hit = cache.get(key)
    if hit is None:
        hit = (build_plan(edge_index, base, stride, K, num_nodes, self_loops), base)   # keep base alive
        cache[key] = hit
    return hit[0], k
