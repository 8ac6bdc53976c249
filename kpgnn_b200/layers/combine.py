"""Combine functions over the hop axis -- mirror of the reference's layers/combine.py.

`from layers.combine import *` must also export `torch` and `nn` (run_simulation.py:23 relies on it).
GeometricCombine (combine.py:30-58) is normally FUSED into the aggregation kernel by the layers that own one
(KPGINPlusConv, KPGCNConv): they call `thetas()` and hand the [K,d] weights to kp_agg_forward, so the [N,K,d]
tensor never exists.  `forward` is kept for callers that combine an existing tensor (KPGINConv, KPGraphSAGEConv
apply it after their per-hop MLP).  AttentionCombine (combine.py:8-27) runs as one kernel per direction on CUDA
tensors (kp_attn_combine_*, csrc/attn.cu: bi-LSTM recurrences, softmax over hops and the weighted sum on one warp per
node; the [N,K,2K] LSTM output and the [N,K,d] product never exist); its parameters stay an `nn.LSTM` module so the
reference's state_dict keys (`attention_lstm.weight_ih_l0` ...) are unchanged.
"""
import torch
import torch.nn as nn


class _AttnCombineFn(torch.autograd.Function):
    """kp_attn_combine_forward / _backward (include/kpgnn.h).  Parameter order: the four forward-direction tensors
    (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0), then the four `_reverse` ones."""

    @staticmethod
    def _desc(x, params):
        from .. import _lib
        N, K, d = x.shape
        desc = _lib.AttnDesc()
        desc.N, desc.K, desc.d = N, K, d
        desc.x, desc.x_node_stride, desc.x_hop_stride = x.data_ptr(), x.stride(0), x.stride(1)
        for dirn in range(2):
            w_ih, w_hh, b_ih, b_hh = params[4 * dirn:4 * dirn + 4]
            desc.w_ih[dirn], desc.w_hh[dirn] = w_ih.data_ptr(), w_hh.data_ptr()
            desc.b_ih[dirn], desc.b_hh[dirn] = b_ih.data_ptr(), b_hh.data_ptr()
        return desc

    @staticmethod
    def forward(ctx, x, *params):
        import ctypes as C
        from .. import _lib
        lib = _lib.lib()
        x = x.detach()
        if x.stride(-1) != 1:
            x = x.contiguous()
        params = tuple(p.detach().contiguous() for p in params)
        N, K, d = x.shape
        out = torch.empty((N, d), dtype=torch.float32, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        desc = _AttnCombineFn._desc(x, params)
        _lib.check(lib.kp_attn_combine_forward(C.byref(desc), out.data_ptr(), None, st), "kp_attn_combine_forward")
        ctx.x, ctx.params = x, params
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        from .. import _lib
        lib = _lib.lib()
        x, params = ctx.x, ctx.params
        N, K, d = x.shape
        dev = x.device
        dout = dout.contiguous()
        desc = _AttnCombineFn._desc(x, params)
        dx = torch.empty((N, K, d), dtype=torch.float32, device=dev)
        dwi = [torch.empty((4 * K, d), dtype=torch.float32, device=dev) for _ in range(2)]
        dwh = [torch.empty((4 * K, K), dtype=torch.float32, device=dev) for _ in range(2)]
        db = [torch.empty(4 * K, dtype=torch.float32, device=dev) for _ in range(2)]
        nb = C.c_size_t(0)
        _lib.check(lib.kp_attn_combine_backward_workspace_bytes(C.byref(desc), C.byref(nb)), "kp_attn_combine ws")
        ws = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.kp_attn_combine_backward(C.byref(desc), dout.data_ptr(), dx.data_ptr(), dwi[0].data_ptr(),
                                                dwi[1].data_ptr(), dwh[0].data_ptr(), dwh[1].data_ptr(),
                                                db[0].data_ptr(), db[1].data_ptr(), ws.data_ptr(), ws.numel(), st),
                   "kp_attn_combine_backward")
        # b_ih and b_hh enter the gates as a sum: both receive the same gradient
        return (dx, dwi[0], dwh[0], db[0], db[0].clone(), dwi[1], dwh[1], db[1], db[1].clone())


class AttentionCombine(nn.Module):
    """Attention combination.  Args: hidden_size (per-hop width), K (hops)."""

    def __init__(self, hidden_size, K):
        super(AttentionCombine, self).__init__()
        self.attention_lstm = nn.LSTM(hidden_size, K, 1, batch_first=True, bidirectional=True, dropout=0.)

    def reset_parameters(self):
        self.attention_lstm.reset_parameters()

    def _kernel_ok(self, x):
        m = self.attention_lstm
        return (x.is_cuda and x.dim() == 3 and x.dtype == torch.float32 and x.size(1) == m.hidden_size
                and 1 <= m.hidden_size <= 16 and 1 <= x.size(2) <= 128 and x.size(2) == m.input_size)

    def forward(self, x):
        m = self.attention_lstm
        if self._kernel_ok(x):
            return _AttnCombineFn.apply(x, m.weight_ih_l0, m.weight_hh_l0, m.bias_ih_l0, m.bias_hh_l0,
                                        m.weight_ih_l0_reverse, m.weight_hh_l0_reverse, m.bias_ih_l0_reverse,
                                        m.bias_hh_l0_reverse)
        # other shapes (sequence length != hidden size never occurs in the reference's layers): library LSTM
        m.flatten_parameters()
        score, _ = m(x)                                                  # N * K * 2K
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
