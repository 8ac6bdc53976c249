"""Combine functions over the hop axis -- mirror of the reference's layers/combine.py.

`from layers.combine import *` must also export `torch` and `nn` (run_simulation.py:23 relies on it).
GeometricCombine (combine.py:30-58) is normally FUSED into the aggregation kernel by the layers that own one
(KPGINPlusConv, KPGCNConv): they call `thetas()` and hand the [K,d] weights to kp_agg_forward, so the [N,K,d]
tensor never exists.  `forward` is kept for callers that combine an existing tensor (KPGINConv, KPGraphSAGEConv
apply it after their per-hop MLP).  AttentionCombine (combine.py:8-27) needs the [N,K,d] tensor as LSTM input
and stays a cuDNN LSTM + softmax + weighted sum.
"""
import torch
import torch.nn as nn


class AttentionCombine(nn.Module):
    """Attention combination.  Args: hidden_size (per-hop width), K (hops)."""

    def __init__(self, hidden_size, K):
        super(AttentionCombine, self).__init__()
        self.attention_lstm = nn.LSTM(hidden_size, K, 1, batch_first=True, bidirectional=True, dropout=0.)

    def reset_parameters(self):
        self.attention_lstm.reset_parameters()

    def forward(self, x):
        self.attention_lstm.flatten_parameters()
        score, _ = self.attention_lstm(x)                               # N * K * 2K
        weight = torch.softmax(score.sum(dim=-1), dim=1).unsqueeze(-1)   # N * K * 1
        return (x * weight).sum(dim=1)


class _GeoTheta(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alphas, K):
        import ctypes as C
        from .. import _lib
        lib = _lib.lib()
        alphas = alphas.detach().contiguous()
        theta = torch.empty((K, alphas.numel()), dtype=torch.float32, device=alphas.device)
        st = C.c_void_p(torch.cuda.current_stream(alphas.device).cuda_stream)
        _lib.check(lib.kp_geometric_theta_forward(alphas.data_ptr(), K, alphas.numel(), theta.data_ptr(), st),
                   "kp_geometric_theta_forward")
        ctx.save_for_backward(alphas, theta)
        ctx.K = K
        return theta

    @staticmethod
    def backward(ctx, dtheta):
        import ctypes as C
        from .. import _lib
        lib = _lib.lib()
        alphas, theta = ctx.saved_tensors
        dtheta = dtheta.contiguous()
        dal = torch.empty_like(alphas)
        st = C.c_void_p(torch.cuda.current_stream(alphas.device).cuda_stream)
        _lib.check(lib.kp_geometric_theta_backward(alphas.data_ptr(), theta.data_ptr(), dtheta.data_ptr(), ctx.K,
                                                   alphas.numel(), dal.data_ptr(), st), "kp_geometric_theta_backward")
        return dal, None


class GeometricCombine(nn.Module):
    """Geometric combination.  Args: K (hops), hidden_size (per-hop width) -- note the argument order."""

    def __init__(self, K, hidden_size):
        super(GeometricCombine, self).__init__()
        self.alphas = nn.Parameter(torch.zeros(hidden_size))
        self.K = K
        self.hidden_size = hidden_size

    def reset_parameters(self):
        nn.init.zeros_(self.alphas)

    def thetas(self):
        """[K, hidden] softmax over hops of alpha*(1-alpha)^k, alpha = sigmoid(alphas) (combine.py:51-58).
        On a CUDA device this is one kernel forward and one backward (kp_geometric_theta_*) instead of the
        reference's ~20 elementwise launches on a [K, hidden] tensor."""
        if self.alphas.is_cuda:
            return _GeoTheta.apply(self.alphas, self.K)
        a = torch.sigmoid(self.alphas)
        hops = torch.arange(self.K, device=a.device, dtype=a.dtype).unsqueeze(-1)
        return torch.softmax(a * (1 - a) ** hops, dim=0)

    def geometric_distribution(self):
        return self.thetas().unsqueeze(0)

    def forward(self, x):
        return (x * self.thetas()).sum(dim=-2)


class GINEPlusCombine(nn.Module):
    """GINE+ combination (combine.py:61-76; unused by every reference model)."""

    def __init__(self, K):
        super(GINEPlusCombine, self).__init__()
        self.K = K
        self.eps = nn.Parameter(torch.zeros(1, K))

    def reset_parameters(self):
        nn.init.zeros_(self.eps)

    def forward(self, x):
        return ((1 + self.eps.unsqueeze(-1)) * x).sum(dim=1)
