"""KP-GNN GCN layer -- mirror of the reference's layers/KPGCN.py:11-126 on the sm_100a aggregation kernels."""
import torch.nn.functional as F  # noqa: F401

from ._base import KHopLayer, make_combine, khop_aggregate, get_plan, ACT_RELU
from .combine import *  # noqa: F401,F403


def degree(index, num_nodes, index_mask):
    """Per-hop in-degree [N,K] of the hop-masked edge list (KPGCN.py:11-25).  The layer itself reads the same
    numbers off the plan's row pointers; this helper is kept for callers that imported it."""
    out = torch.zeros((num_nodes, index_mask.size(-1)), device=index.device)
    return out.index_add_(0, index, (index_mask > 0).to(out.dtype))


class KPGCNConv(KHopLayer):
    """KP-GNN with GCN kernel.
    Args: input_size, output_size, K, num_hop1_edge, num_pe, combine.
    forward(x [N,in], ...) -> [N, output_size]: self loops (attr 1 in every hop), per-hop symmetric
    normalisation, ReLU, + P, combine, projection.  Loops and degrees live in the plan; with the geometric
    combine  sum_k theta_k (relu(Agg_k) + P_k)  is one kernel.
    """

    def __init__(self, input_size, output_size, K, num_hop1_edge=1, num_pe=1, combine="geometric"):
        super(KPGCNConv, self).__init__()
        self.aggr = "add"
        self.K = K
        self.output_size = output_size
        assert output_size % K == 0
        self.output_dk = output_size // K
        self.hop_proj = nn.Linear(input_size, output_size)
        self.hop1_edge_emb = torch.nn.Embedding(num_hop1_edge + 2, self.output_dk, padding_idx=0)
        if self.K > 1:
            self.hopk_edge_emb = torch.nn.Embedding(num_pe + 2, self.output_dk, padding_idx=0)
            self.hopk_node_path_emb = torch.nn.Embedding(num_pe, self.output_dk, padding_idx=0)
            self.combine_proj = nn.Linear(self.output_dk, output_size)
            self.combine = make_combine(combine, self.K, self.output_dk)
        else:
            self.hopk_edge_emb = None
            self.combine = torch.squeeze
            self.combine_proj = nn.Identity()
        self.reset_parameters()

    def reset_parameters(self):
        self.hop1_edge_emb.reset_parameters()
        self.hop_proj.reset_parameters()
        if self.K > 1:
            self.hopk_edge_emb.reset_parameters()
            self.hopk_node_path_emb.reset_parameters()
            self.combine.reset_parameters()
        if isinstance(self.combine_proj, nn.Linear):
            self.combine_proj.reset_parameters()

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        self._check_hops(edge_attr)
        plan, k = get_plan(edge_index, edge_attr, x.size(0), self_loops=True)
        x = self.hop_proj(x).view(-1, self.K, self.output_dk)
        x = self._add_path_encoding(x, pe_attr)
        t0, tk = self._tables()
        if isinstance(self.combine, GeometricCombine):
            h = khop_aggregate(x, plan, k, P=peripheral_attr, T0=t0, Tk=tk, theta=self.combine.thetas(),
                               act=ACT_RELU, fuse=True, use_dinv=True)
        else:
            h = self.combine(khop_aggregate(x, plan, k, P=peripheral_attr, T0=t0, Tk=tk, act=ACT_RELU,
                                            use_dinv=True))
        return self.combine_proj(h)
