"""K-hop GIN layer of the regular-graph simulation -- mirror of `KGINConv` in the reference's run_simulation.py:29-93
(BASELINE.json config 5: node-level KP-GIN, 3-regular graphs, n up to 1280, K up to 6, hidden 16, forward only).

Same constructor/parameters/state_dict keys; the masked per-hop aggregation (run_simulation.py:73,87-90: no edge
embeddings, `x_j` masked by the hop attr) plus the `(1+eps) x` self term is one call of the no-table mode of the
aggregation kernel.  `graph=True` reproduces the script's `--graph` flag (global_add_pool of the node outputs).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import khop_aggregate, ACT_NONE
from .plan import get_plan


class KGINConv(nn.Module):
    def __init__(self, hidden_size, K, eps=0., train_eps=False, graph=False):
        super(KGINConv, self).__init__()
        self.aggr = "add"
        self.K = K
        self.hidden_size = hidden_size
        self.graph = graph
        self.proj = nn.Linear(1, K * hidden_size)
        self.hop_proj1 = torch.nn.Parameter(torch.Tensor(self.K, hidden_size, hidden_size))
        self.hop_bias1 = torch.nn.Parameter(torch.Tensor(self.K, hidden_size))
        self.hop_proj2 = torch.nn.Parameter(torch.Tensor(self.K, hidden_size, hidden_size))
        self.hop_bias2 = torch.nn.Parameter(torch.Tensor(self.K, hidden_size))
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        self.combine_proj = nn.Linear(hidden_size * K, hidden_size)
        self.reset_parameters()

    def reset_parameters(self):
        for w, b in ((self.hop_proj1, self.hop_bias1), (self.hop_proj2, self.hop_bias2)):
            nn.init.kaiming_uniform_(w)
        for w, b in ((self.hop_proj1, self.hop_bias1), (self.hop_proj2, self.hop_bias2)):
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(w)
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(b, -bound, bound)
        self.combine_proj.reset_parameters()
        nn.init.zeros_(self.eps)

    def forward(self, x, edge_index, edge_attr, batch):
        x = self.proj(x).view(-1, self.K, self.hidden_size)
        plan, k = get_plan(edge_index, edge_attr, x.size(0))
        z = khop_aggregate(x, plan, k, eps=self.eps, act=ACT_NONE)          # Agg + (1+eps) x
        z = z.permute(1, 0, 2)
        z = F.relu(torch.matmul(z, self.hop_proj1) + self.hop_bias1.unsqueeze(1))
        z = F.relu(torch.matmul(z, self.hop_proj2) + self.hop_bias2.unsqueeze(1))
        z = z.permute(1, 0, 2).contiguous().view(-1, self.K * self.hidden_size)
        out = self.combine_proj(z)
        if self.graph:
            ng = int(batch.max()) + 1
            out = torch.zeros((ng, out.size(1)), dtype=out.dtype, device=out.device).index_add_(0, batch, out)
        return out
