"""KP-GNN GIN layer -- mirror of the reference's layers/KPGIN.py:12-121 on the sm_100a aggregation kernels."""
import math

import torch.nn.functional as F

from ._base import KHopLayer, make_combine, khop_aggregate, get_plan, ACT_NONE
from .combine import *  # noqa: F401,F403


class KPGINConv(KHopLayer):
    """KP-GNN with GIN kernel.
    Args: input_size, output_size, K, eps (float), train_eps (bool), num_hop1_edge, num_pe, combine.
    forward(x [N,H], edge_index, edge_attr [E,K], pe_attr, peripheral_attr [N,K,dk]) -> [N, output_size].
    One kernel produces Z = Agg + P + (1+eps) x  as [N,K,dk]; the per-hop 2-layer MLP stays two batched GEMMs.
    """

    def __init__(self, input_size, output_size, K, eps=0., train_eps=False, num_hop1_edge=1, num_pe=1,
                 combine="geometric"):
        super(KPGINConv, self).__init__()
        self.aggr = "add"
        self.K = K
        self.output_size = output_size
        assert input_size % K == 0
        assert output_size % K == 0
        self.input_dk = input_size // K
        self.output_dk = output_size // K
        self.hop_proj1 = torch.nn.Parameter(torch.Tensor(self.K, self.input_dk, self.output_dk))
        self.hop_bias1 = torch.nn.Parameter(torch.Tensor(self.K, self.output_dk))
        self.hop_proj2 = torch.nn.Parameter(torch.Tensor(self.K, self.output_dk, self.output_dk))
        self.hop_bias2 = torch.nn.Parameter(torch.Tensor(self.K, self.output_dk))
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        # +2 rows: 0 = mask, 1 = self connection
        self.hop1_edge_emb = torch.nn.Embedding(num_hop1_edge + 2, self.input_dk, padding_idx=0)
        if self.K > 1:
            self.hopk_edge_emb = torch.nn.Embedding(num_pe + 2, self.input_dk, padding_idx=0)
            self.hopk_node_path_emb = torch.nn.Embedding(num_pe, self.input_dk, padding_idx=0)
            self.combine_proj = nn.Linear(self.output_dk, output_size)
            self.combine = make_combine(combine, self.K, self.output_dk)
        else:
            self.hopk_edge_emb = None
            self.combine = torch.squeeze
            self.combine_proj = nn.Identity()
        self.reset_parameters()

    def reset_parameters(self):
        self.hop1_edge_emb.reset_parameters()
        for w, b in ((self.hop_proj1, self.hop_bias1), (self.hop_proj2, self.hop_bias2)):
            nn.init.kaiming_uniform_(w)
        for w, b in ((self.hop_proj1, self.hop_bias1), (self.hop_proj2, self.hop_bias2)):
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(w)
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(b, -bound, bound)
        if self.K > 1:
            self.hopk_edge_emb.reset_parameters()
            self.hopk_node_path_emb.reset_parameters()
            self.combine.reset_parameters()
        if isinstance(self.combine_proj, nn.Linear):
            self.combine_proj.reset_parameters()
        nn.init.zeros_(self.eps)

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        self._check_hops(edge_attr)
        x = x.view(-1, self.K, self.input_dk)
        plan, k = get_plan(edge_index, edge_attr, x.size(0))
        x = self._add_path_encoding(x, pe_attr)
        t0, tk = self._tables()
        z = khop_aggregate(x, plan, k, P=peripheral_attr, T0=t0, Tk=tk, eps=self.eps, act=ACT_NONE)
        z = z.permute(1, 0, 2)
        z = F.relu(torch.matmul(z, self.hop_proj1) + self.hop_bias1.unsqueeze(1))
        z = F.relu(torch.matmul(z, self.hop_proj2) + self.hop_bias2.unsqueeze(1))
        z = z.permute(1, 0, 2)
        return self.combine_proj(self.combine(z))
