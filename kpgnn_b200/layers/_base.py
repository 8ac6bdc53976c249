"""Shared plumbing of the drop-in K-hop layers (not part of the reference's API surface)."""
import torch
import torch.nn as nn

from ..ops import khop_aggregate, ACT_NONE, ACT_GELU, ACT_RELU  # noqa: F401
from ..plan import get_plan
from .combine import AttentionCombine, GeometricCombine


def make_combine(kind, K, width):
    if kind == "attention":
        return AttentionCombine(width, K)
    if kind == "geometric":
        return GeometricCombine(K, width)
    raise ValueError("Not implemented combine function")


def _all_zero(t):
    """True iff the integer tensor `t` is identically zero; cached on its base tensor (one sync per batch)."""
    base = t._base if t._base is not None else t
    tag = getattr(base, "_kpgnn_allzero", None)
    if tag is None or tag[0] != base._version:
        tag = (base._version, not bool((base != 0).any()))
        try:
            base._kpgnn_allzero = tag
        except Exception:  # pragma: no cover
            pass
    return tag[1]


class KHopLayer(nn.Module):
    """Base: embedding tables + the reference's in-place path-encoding add."""

    def _tables(self):
        t0 = self.hop1_edge_emb.weight
        tk = self.hopk_edge_emb.weight if self.hopk_edge_emb is not None else None
        return t0, tk

    def _add_path_encoding(self, x, pe_attr):
        """x[:, 1:] += hopk_node_path_emb(pe_attr), in place on the caller's storage like the reference
        (KPGIN.py:92-94).  The reference's extractor always emits pe_attr == 0 (data_utils.py:91 reads a diagonal
        that adj_K_order zeroed, :123) and row 0 of the table is the zero padding row, so the add is skipped when
        pe_attr is identically zero; any other input takes the reference's own torch ops."""
        if self.K > 1 and pe_attr is not None and pe_attr.numel() > 0 and not _all_zero(pe_attr):
            x[:, 1:] = x[:, 1:] + self.hopk_node_path_emb(pe_attr)
        return x

    def _check_hops(self, edge_attr):
        k = edge_attr.size(1) if edge_attr.dim() == 2 else 1
        if k != self.K:
            raise ValueError("edge_attr has %d hop columns but the layer was built with K=%d" % (k, self.K))


class _SplitKLinearFn(torch.autograd.Function):
    """y = x W^T + b with a weight gradient computed as a batched split-K GEMM.  The dense GEMMs stay library
    calls (cuBLAS fp32); the only change is the SHAPE handed to the library: dW = dY^T X with N rows of a few
    thousand and a 104x104 output is a 4-CTA launch for cuBLAS' fp32 kernels (26-42 us measured, profiles/),
    while S independent [out x N/S] x [N/S x in] products fill the machine."""
    SPLIT = 32

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = dy @ weight
        if ctx.needs_input_grad[1]:
            n, S = x.size(0), _SplitKLinearFn.SPLIT
            m = n // S
            if m >= 16:
                main = m * S
                dw = torch.bmm(dy[:main].view(S, m, -1).transpose(1, 2), x[:main].view(S, m, -1)).sum(0)
                if main < n:
                    dw = dw + dy[main:].t() @ x[main:]
            else:
                dw = dy.t() @ x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


class SplitKLinear(nn.Linear):
    """nn.Linear (same parameters / state_dict keys) whose weight gradient is a split-K batched GEMM."""

    def forward(self, x):
        if x.dim() == 2 and x.is_cuda:
            return _SplitKLinearFn.apply(x, self.weight, self.bias)
        return super().forward(x)
