"""Autograd binding of the fused K-hop aggregation kernels (kp_agg_forward / kp_agg_backward, include/kpgnn.h).

`khop_aggregate` is the single operator all five drop-in layers call in place of PyG's
`self.propagate(edge_index, x=..., edge_emb=..., mask=edge_attr)` + update + (optionally) GeometricCombine
(layers/KPGIN.py:100-105, KPGINplus.py:74-78, KPGCN.py:107-116, KPGraphSAGE.py:86-89, gine.py:52-53).
PyTorch only provides device memory, the current stream and the autograd graph here.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU  # noqa: F401


# average entries per (node, hop) row from which the block-resident kernels (csrc/agg_tile.cu) are worth their staging
LONG_ROW_ENTRIES = 12
# Size of the backward's hand-over tensor Gs [N,k,d] (bytes) from which the whole backward runs as ONE block-resident
# kernel (csrc/agg_block_bwd.cu: Gs stays in shared memory, DRAM traffic = the algorithmic bytes).  OFF by default
# (None): measured at 8 192 molecules the fused kernel takes 2.30 ms against 1.10 ms for the three pipelined kernels --
# a molecule gives a 512-thread CTA only ~23 rows per phase, and its three barrier-separated phases expose the
# rowptr -> entries -> gather latency chain that the lean kernels hide by software pipelining across nodes
# (profiles/r2_block_bwd.txt).  Parity-tested (tests/test_tile_gpu.py); set a byte threshold to opt in.
BLOCK_BWD_MIN_BYTES = None


# Node-range (chunked) backward, include/kpgnn.h kp_agg_desc.node_base: the backward of a batch whose hand-over tensor
# Gs [N,k,d] is larger than the L2 run over ranges of whole graphs, every chunk reusing ONE slice-sized Gs workspace that
# stays in L2 between B1, B2 and B3.  Parity-tested (tests/test_chunked_bwd_gpu.py) but OFF by default
# (CHUNK_BWD_MIN_BYTES = None): measured at 8 192 molecules (profiles/r2_chunked_bwd.txt) the single call takes 1.11 ms
# and the chunked one 1.37 ms (7 chunks) to 1.66 ms (26 chunks) of pure device time -- each chunk pays the persistent
# kernels' fixed costs (table staging, accumulator reset, partial write-out, a ragged last wave) five launches over,
# which is more than the two saved HBM passes over Gs.  Set a byte threshold to opt in.
CHUNK_BWD_MIN_BYTES = None
CHUNK_BWD_GS_BYTES = int(os.environ.get("KP_CHUNK_GS_MB", "92")) << 20
if os.environ.get("KP_CHUNK_BWD") == "1":
    CHUNK_BWD_MIN_BYTES = 192 << 20


def backward_chunks(plan, k, d):
    """Node boundaries [0, ..., N] at graph (closed-block) boundaries for the chunked backward, or None."""
    if CHUNK_BWD_MIN_BYTES is None or 4 * plan.N * k * d < CHUNK_BWD_MIN_BYTES:
        return None
    bp = plan.block_ptr_host()
    if bp is None or len(bp) < 3:
        return None
    per = max(1, CHUNK_BWD_GS_BYTES // (4 * k * d))
    bounds, N = [0], int(bp[-1])
    import numpy as np
    while bounds[-1] < N:
        j = int(np.searchsorted(bp, bounds[-1] + per, side="right")) - 1
        nxt = int(bp[j])
        if nxt <= bounds[-1]:
            nxt = int(bp[j + 1])                  # a single block larger than the target: take it whole
        bounds.append(nxt)
    return bounds if len(bounds) > 2 else None


def agg_backward(plan, desc, dout, dX, dP, dT0, dTk, dth, deps, stream=None, chunks=None):
    """kp_agg_backward for one call, over node ranges when `chunks` (backward_chunks) is given and the kernels that
    would run honour it; table / theta / alpha gradients of the chunks are added in chunk order (deterministic).
    `desc` is this call's descriptor (not modified); `dP` is the fused call's dP or None."""
    lib = _lib.lib()
    dev = dout.device
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream)
    if chunks is not None and deps is None:
        probe = _lib.AggDesc.from_buffer_copy(desc)
        probe.block_ptr, probe.block_stats, probe.num_blocks, probe.max_block_nodes = None, None, 0, 0
        ok = C.c_int32(0)
        _lib.check(lib.kp_agg_backward_chunkable(C.byref(probe), C.byref(ok)), "kp_agg_backward_chunkable")
        if not ok.value:
            chunks = None
    else:
        chunks = None
    if chunks is None:
        nbytes = C.c_size_t(0)
        _lib.check(lib.kp_agg_backward_workspace_bytes(C.byref(desc), C.byref(nbytes)), "kp_agg_backward_workspace_bytes")
        ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        _lib.check(lib.kp_agg_backward(C.byref(desc), dout.data_ptr(), _ptr(dX), _ptr(dP), _ptr(dT0), _ptr(dTk), _ptr(dth),
                                       _ptr(deps), ws.data_ptr(), ws.numel(), st), "kp_agg_backward")
        return ws
    k, d, Kp, fuse = desc.k, desc.d, desc.Kplan, bool(desc.fuse)
    geo = bool(desc.geo_alphas) and bool(desc.geo_dalphas)
    descs, ws_bytes = [], 0
    for n0, n1 in zip(chunks[:-1], chunks[1:]):
        dc = _lib.AggDesc.from_buffer_copy(probe)
        dc.N, dc.node_base, dc.leaf_stream = n1 - n0, n0, None
        dc.rowptr, dc.rowptrT = desc.rowptr + 4 * n0 * Kp, desc.rowptrT + 4 * n0 * Kp
        if desc.P:
            dc.P = desc.P + 4 * n0 * (desc.p_node_stride or k * d)
        nb = C.c_size_t(0)
        _lib.check(lib.kp_agg_backward_workspace_bytes(C.byref(dc), C.byref(nb)), "kp_agg_backward_workspace_bytes")
        ws_bytes = max(ws_bytes, nb.value)
        descs.append(dc)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    nc = len(descs)
    pT0 = torch.empty((nc,) + tuple(dT0.shape), dtype=torch.float32, device=dev) if dT0 is not None else None
    pTk = torch.empty((nc,) + tuple(dTk.shape), dtype=torch.float32, device=dev) if dTk is not None else None
    pth = torch.empty((nc, k, d), dtype=torch.float32, device=dev) if dth is not None else None
    pal = torch.empty((nc, d), dtype=torch.float32, device=dev) if geo else None
    do_row = d if fuse else k * d
    dx_row = desc.dx_node_stride or k * d
    for i, dc in enumerate(descs):
        n0 = dc.node_base
        if geo:
            dc.geo_dalphas = pal[i].data_ptr()
        _lib.check(lib.kp_agg_backward(
            C.byref(dc), dout.data_ptr() + 4 * n0 * do_row, None if dX is None else dX.data_ptr() + 4 * n0 * dx_row,
            None if dP is None else dP.data_ptr() + 4 * n0 * k * d, None if pT0 is None else pT0[i].data_ptr(),
            None if pTk is None else pTk[i].data_ptr(), None if pth is None else pth[i].data_ptr(), None,
            ws.data_ptr(), ws.numel(), st), "kp_agg_backward (chunk %d)" % i)
    ext = torch.cuda.ExternalStream(st.value, device=dev) if stream is not None else None
    with torch.cuda.stream(ext) if ext is not None else _nullctx():
        if pT0 is not None:
            torch.sum(pT0, dim=0, out=dT0)
        if pTk is not None:
            torch.sum(pTk, dim=0, out=dTk)
        if pth is not None:
            torch.sum(pth, dim=0, out=dth)
        if geo:
            C_f = C.c_void_p(desc.geo_dalphas)
            out = torch.as_tensor(_RawF32(C_f.value, d), device=dev)
            torch.sum(pal, dim=0, out=out)
    return ws


class _RawF32(object):
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class _nullctx(object):
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def want_blocks(plan, k, d, fuse):
    """Asks the plan for its closed node blocks when a block-resident kernel would serve this call."""
    if plan.block_ptr is not None:
        return
    if not fuse and plan.nnz >= LONG_ROW_ENTRIES * max(plan.N * plan.K, 1):
        plan.blocks()          # long rows (e.g. n = 1 280 regular graphs at K = 6): stage blocks in shared memory
    elif BLOCK_BWD_MIN_BYTES is not None and 4 * plan.N * k * d >= BLOCK_BWD_MIN_BYTES:
        plan.blocks()          # opt-in: keep the backward's hand-over tensor in shared memory


def _ptr(t):
    return None if t is None else t.data_ptr()


def _make_desc(plan, k, x, P, T0, Tk, theta, eps, act, fuse, use_dinv, use_mean):
    N, kk, d = x.shape
    assert kk == k and N == plan.N, (x.shape, k, plan.N)
    desc = _lib.AggDesc()
    desc.N, desc.Kplan, desc.k, desc.d = N, plan.K, k, d
    desc.rowptr, desc.col, desc.attr16 = plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.attr16.data_ptr()
    desc.rowptrT, desc.colT = plan.rowptrT.data_ptr(), plan.colT.data_ptr()
    desc.dinv = plan.dinv.data_ptr() if use_dinv else None
    desc.indeg = plan.indeg.data_ptr() if use_mean else None
    desc.X, desc.x_node_stride, desc.x_hop_stride = x.data_ptr(), x.stride(0), x.stride(1)
    if P is not None:
        desc.P, desc.p_node_stride, desc.p_hop_stride = P.data_ptr(), P.stride(0), P.stride(1)
    else:
        desc.P, desc.p_node_stride, desc.p_hop_stride = None, 0, 0
    desc.T0, desc.Tk = _ptr(T0), _ptr(Tk)
    desc.rows0 = T0.size(0) if T0 is not None else 0
    desc.rowsk = Tk.size(0) if Tk is not None else 0
    desc.theta, desc.eps = _ptr(theta), _ptr(eps)
    desc.act, desc.fuse = act, 1 if fuse else 0
    desc.amax0, desc.amaxk = plan.max_attr0, plan.max_attrk
    if plan.block_ptr is not None:
        # closed node blocks were requested for this plan: block-resident kernels where they fit (long rows: unfused
        # forward / dX from shared-memory slices; molecule batches: the whole backward as one kernel)
        desc.block_ptr, desc.block_stats = plan.block_ptr.data_ptr(), plan.block_stats.data_ptr()
        desc.num_blocks, desc.max_block_nodes = plan.num_blocks, plan.max_block_nodes
    return desc


def _prep(t, last_contig=True):
    """fp32 CUDA tensor whose last dim is dense; other strides are passed to the kernel as they are."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError("kpgnn_b200 kernels are fp32 (got %s)" % t.dtype)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


class _KHopAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, P, T0, Tk, theta, eps, plan, k, act, fuse, use_dinv, use_mean):
        lib = _lib.lib()
        if not x.is_cuda:
            raise _lib.KpError("kpgnn_b200 runs on CUDA tensors only (no CPU fallback); got x on %s" % x.device)
        x = _prep(x.detach())
        P_ = _prep(P.detach()) if P is not None else None
        T0_ = T0.detach().contiguous() if T0 is not None else None
        Tk_ = Tk.detach().contiguous() if Tk is not None else None
        th_ = theta.detach().contiguous() if theta is not None else None
        eps_ = eps.detach().contiguous() if eps is not None else None
        N, _, d = x.shape
        if T0_ is not None:
            plan.check_tables(T0_.size(0), Tk_.size(0) if Tk_ is not None else 0, k)
        desc = _make_desc(plan, k, x, P_, T0_, Tk_, th_, eps_, act, fuse, use_dinv, use_mean)
        out = torch.empty((N, d) if fuse else (N, k, d), dtype=torch.float32, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.kp_agg_forward(C.byref(desc), out.data_ptr(), st), "kp_agg_forward")
        ctx.plan, ctx.cfg = plan, (k, act, fuse, use_dinv, use_mean)
        # saved by hand (not save_for_backward): x may be a slot of a caller-managed ring buffer whose OTHER
        # slots are written in place later; version-counter checks on the shared base would be false alarms.
        ctx.saved = (x, P_, T0_, Tk_, th_, eps_)
        ctx.p_is_none = P is None
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        x, P_, T0_, Tk_, th_, eps_ = ctx.saved
        plan = ctx.plan
        k, act, fuse, use_dinv, use_mean = ctx.cfg
        need = ctx.needs_input_grad
        N, _, d = x.shape
        dout = dout.contiguous()
        if dout.data_ptr() % 16:
            dout = dout.clone()
        desc = _make_desc(plan, k, x, P_, T0_, Tk_, th_, eps_, act, fuse, use_dinv, use_mean)
        dev = x.device
        dX = torch.empty((N, k, d), dtype=torch.float32, device=dev) if need[0] else None
        dP = None
        if need[1] and P_ is not None:
            dP = dout if not fuse else torch.empty((N, k, d), dtype=torch.float32, device=dev)
        dT0 = torch.empty_like(T0_) if (T0_ is not None and need[2]) else None
        dTk = torch.empty_like(Tk_) if (Tk_ is not None and need[3]) else None
        if (dT0 is None) != (dTk is None) and Tk_ is not None:
            # the kernel produces both tables in one pass; allocate the unwanted one too
            dT0 = dT0 if dT0 is not None else torch.empty_like(T0_)
            dTk = dTk if dTk is not None else torch.empty_like(Tk_)
        dth = torch.empty_like(th_) if (th_ is not None and need[4] and fuse) else None
        deps = torch.empty_like(eps_) if (eps_ is not None and need[5]) else None
        agg_backward(plan, desc, dout, dX, dP if fuse else None, dT0, dTk, dth, deps,
                     chunks=backward_chunks(plan, k, d))
        return (dX, dP, dT0 if need[2] else None, dTk if need[3] else None, dth, deps,
                None, None, None, None, None, None)


def khop_aggregate(x, plan, k, P=None, T0=None, Tk=None, theta=None, eps=None, act=ACT_NONE, fuse=False,
                   use_dinv=False, use_mean=False):
    """x [N,k,d] fp32 (any node/hop strides, dense last dim) -> [N,d] if fuse else [N,k,d].  See kpgnn.h."""
    if fuse and theta is None:
        raise ValueError("fuse=True needs theta [k,d]")
    want_blocks(plan, k, x.size(-1), fuse)
    return _KHopAggregate.apply(x, P, T0, Tk, theta, eps, plan, k, act, fuse, use_dinv, use_mean)
