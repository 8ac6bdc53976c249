"""KP-GNN GraphSAGE layer -- mirror of the reference's layers/KPGraphSAGE.py:12-106 on the sm_100a kernels."""
import math

import torch.nn.functional as F

from ._base import KHopLayer, make_combine, khop_aggregate, get_plan, ACT_NONE
from .combine import *  # noqa: F401,F403


class KPGraphSAGEConv(KHopLayer):
    """KP-GNN with GraphSAGE kernel.
    Args: input_size, output_size, K, aggr ("add" is what every reference script passes; "mean" divides by the
    node's K-hop in-edge count irrespective of hop masks, PyG scatter-mean semantics), num_hop1_edge, num_pe, combine.
    """

    def __init__(self, input_size, output_size, K, aggr="mean", num_hop1_edge=1, num_pe=1, combine="geometric"):
        super(KPGraphSAGEConv, self).__init__()
        if aggr not in ("add", "sum", "mean"):
            raise NotImplementedError("aggr=%r (supported: add, mean)" % (aggr,))
        self.aggr = aggr
        self.K = K
        assert input_size % K == 0
        assert output_size % K == 0
        self.input_dk = input_size // K
        self.output_dk = output_size // K
        self.output_size = output_size
        self.hop_proj = torch.nn.Parameter(torch.Tensor(self.K, 2 * self.input_dk, self.output_dk))
        self.hop_bias = torch.nn.Parameter(torch.Tensor(self.K, self.output_dk))
        self.hop1_edge_emb = torch.nn.Embedding(num_hop1_edge + 2, self.input_dk, padding_idx=0)
        if self.K > 1:
            self.combine_proj = nn.Linear(self.output_dk, output_size)
            self.hopk_edge_emb = torch.nn.Embedding(num_pe + 2, self.input_dk, padding_idx=0)
            self.hopk_node_path_emb = torch.nn.Embedding(num_pe, self.input_dk, padding_idx=0)
            self.combine = make_combine(combine, self.K, self.output_dk)
        else:
            self.hopk_edge_emb = None
            self.combine = torch.squeeze
            self.combine_proj = nn.Identity()
        self.reset_parameters()

    def reset_parameters(self):
        self.hop1_edge_emb.reset_parameters()
        if self.K > 1:
            self.hopk_edge_emb.reset_parameters()
            self.hopk_node_path_emb.reset_parameters()
            self.combine.reset_parameters()
        nn.init.kaiming_uniform_(self.hop_proj)
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.hop_proj)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.hop_bias, -bound, bound)
        if isinstance(self.combine_proj, nn.Linear):
            self.combine_proj.reset_parameters()

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        self._check_hops(edge_attr)
        x = x.view(-1, self.K, self.input_dk)
        plan, k = get_plan(edge_index, edge_attr, x.size(0))
        x = self._add_path_encoding(x, pe_attr)
        t0, tk = self._tables()
        x_n = khop_aggregate(x, plan, k, P=peripheral_attr, T0=t0, Tk=tk, act=ACT_NONE,
                             use_mean=(self.aggr == "mean"))
        y = torch.cat([x, x_n], dim=-1).permute(1, 0, 2)
        y = (torch.matmul(y, self.hop_proj) + self.hop_bias.unsqueeze(1)).permute(1, 0, 2)
        y = F.normalize(F.relu(y), p=2, dim=-1)
        return self.combine_proj(self.combine(y))
