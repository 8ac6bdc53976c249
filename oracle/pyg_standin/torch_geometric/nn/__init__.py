"""Stand-in for torch_geometric.nn (test infrastructure): MessagePassing + the norms/pools models/GNNs.py names."""
import inspect

import torch
import torch.nn as tnn


class MessagePassing(tnn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        assert flow == "source_to_target"

    def propagate(self, edge_index, size=None, **kwargs):
        assert self.node_dim == 0
        src, dst = edge_index[0], edge_index[1]
        x = kwargs.get("x", None)
        n = x.size(0) if size is None else size
        params = list(inspect.signature(self.message).parameters.keys())
        margs = {}
        for p in params:
            if p.endswith("_j"):
                margs[p] = kwargs[p[:-2]].index_select(0, src)
            elif p.endswith("_i"):
                margs[p] = kwargs[p[:-2]].index_select(0, dst)
            else:
                margs[p] = kwargs[p]
        msg = self.message(**margs)
        out = self.aggregate(msg, dst, n)
        return self.update(out)

    def aggregate(self, msg, index, dim_size):
        shape = (dim_size,) + tuple(msg.shape[1:])
        if self.aggr in ("add", "sum"):
            return torch.zeros(shape, dtype=msg.dtype, device=msg.device).index_add_(0, index, msg)
        if self.aggr == "mean":
            out = torch.zeros(shape, dtype=msg.dtype, device=msg.device).index_add_(0, index, msg)
            cnt = torch.zeros(dim_size, dtype=msg.dtype, device=msg.device).index_add_(
                0, index, torch.ones_like(index, dtype=msg.dtype)).clamp_(min=1)
            return out / cnt.view((-1,) + (1,) * (msg.dim() - 1))
        raise NotImplementedError(self.aggr)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


def global_add_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return torch.zeros((size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    s = global_add_pool(x, batch, size)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return s / cnt.clamp_(min=1).view(-1, 1)


def global_max_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    out = torch.full((size,) + tuple(x.shape[1:]), float("-inf"), dtype=x.dtype, device=x.device)
    return out.scatter_reduce(0, batch.view(-1, 1).expand_as(x), x, reduce="amax")


class BatchNorm(tnn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = tnn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x):
        return self.module(x)


class LayerNorm(tnn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.eps = eps
        self.weight = tnn.Parameter(torch.ones(in_channels))
        self.bias = tnn.Parameter(torch.zeros(in_channels))

    def reset_parameters(self):
        tnn.init.ones_(self.weight)
        tnn.init.zeros_(self.bias)

    def forward(self, x, batch=None):
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        return out * self.weight + self.bias


class InstanceNorm(tnn.InstanceNorm1d):
    def forward(self, x, batch=None):
        return super().forward(x.t().unsqueeze(0)).squeeze(0).t()


class PairNorm(tnn.Module):
    def __init__(self, scale=1., scale_individually=False, eps=1e-5):
        super().__init__()
        self.scale, self.eps = scale, eps

    def forward(self, x, batch=None):
        x = x - x.mean(dim=0, keepdim=True)
        return self.scale * x / (self.eps + x.pow(2).sum(-1).mean()).sqrt()


class GraphSizeNorm(tnn.Module):
    def forward(self, x, batch=None):
        if batch is None:
            batch = torch.zeros(x.size(0), dtype=torch.long, device=x.device)
        deg = torch.zeros(int(batch.max()) + 1, dtype=x.dtype, device=x.device).index_add_(
            0, batch, torch.ones_like(batch, dtype=x.dtype))
        return x * deg.pow(-0.5)[batch].view(-1, 1)


class AttentionalAggregation(tnn.Module):
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn, self.nn = gate_nn, nn

    def reset_parameters(self):
        for m in (self.gate_nn, self.nn):
            if m is not None and hasattr(m, "reset_parameters"):
                m.reset_parameters()

    def forward(self, x, index, dim_size=None):
        size = int(index.max()) + 1 if dim_size is None else dim_size
        gate = self.gate_nn(x).view(-1, 1)
        x = self.nn(x) if self.nn is not None else x
        gmax = torch.full((size, 1), float("-inf"), dtype=x.dtype).scatter_reduce(0, index.view(-1, 1), gate, "amax")
        e = (gate - gmax[index]).exp()
        den = torch.zeros(size, 1, dtype=x.dtype).index_add_(0, index, e)
        return global_add_pool(e / den[index] * x, index, size)


class DataParallel(tnn.Module):
    def __init__(self, module, device_ids=None, **kw):
        super().__init__()
        self.module = module

    def forward(self, data_list):
        from ..data import Batch
        return self.module(Batch.from_data_list(data_list))
