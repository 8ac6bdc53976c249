"""GINE layer -- mirror of the reference's layers/gine.py:9-59 on the sm_100a aggregation kernels.

GNNPrime (models/GNNs.py:679) calls this with the full K-hop `edge_index` and `edge_attr[:, :1]`; the plan of
the base [E,K] tensor already holds the hop-1 rows compacted, so the >90% of K-hop edges whose hop-1 attr is 0
are never touched (the reference gathers and masks all of them).
"""
import torch
import torch.nn as nn

from ..ops import khop_aggregate, ACT_NONE
from ..plan import get_plan
from ._base import SplitKLinear
from .norm import FusedBatchNorm1d


class GINEConv(nn.Module):
    """Args: input_size, output_size, eps, num_hop1_edge, train_eps."""

    def __init__(self, input_size, output_size, eps=0., num_hop1_edge=1, train_eps=False):
        super(GINEConv, self).__init__()
        self.input_size = input_size
        self.output_size = output_size
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        self.mlp = nn.Sequential(   # Linear-BN-ReLU x2; the ReLUs are folded into the BatchNorm kernels
            SplitKLinear(input_size, output_size), FusedBatchNorm1d(output_size, relu=True), nn.Identity(),
            SplitKLinear(output_size, output_size), FusedBatchNorm1d(output_size, relu=True), nn.Identity())
        self.hop1_edge_emb = torch.nn.Embedding(num_hop1_edge + 2, self.input_size, padding_idx=0)
        self.reset_parameters()

    def weights_init(self, m):
        if hasattr(m, "reset_parameters"):
            m.reset_parameters()

    def reset_parameters(self):
        self.mlp.apply(self.weights_init)
        self.hop1_edge_emb.reset_parameters()
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index, edge_attr):
        x = x.view(-1, 1, self.input_size)
        plan, k = get_plan(edge_index, edge_attr, x.size(0))
        if k != 1:
            raise ValueError("GINEConv expects edge_attr with one column, got %d" % k)
        out = khop_aggregate(x, plan, 1, T0=self.hop1_edge_emb.weight, eps=self.eps, act=ACT_NONE)
        return self.mlp(out.squeeze())
