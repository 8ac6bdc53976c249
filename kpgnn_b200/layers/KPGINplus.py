"""KP-GIN+ layer -- mirror of the reference's layers/KPGINplus.py:10-88 on the fused sm_100a aggregation."""
import torch.nn.functional as F  # noqa: F401  (star-import surface parity with the reference module)

from ._base import KHopLayer, SplitKLinear, make_combine, khop_aggregate, get_plan, ACT_GELU
from .norm import FusedBatchNorm1d
from .dense_block import fused_dense_block
from .combine import *  # noqa: F401,F403


class KPGINPlusConv(KHopLayer):
    """KP-GNN with GIN plus convolution kernel.
    Args:
        input_size (int), output_size (int), K (int): hops, num_hop1_edge (int), num_pe (int),
        combine (str): geometric | attention (the default "independent" raises for K>1, as in the reference)
    forward(x [N,K,H], edge_index, edge_attr [E,K], pe_attr, peripheral_attr [N,K,H]) -> [N, output_size]:
        mlp( sum_k theta_k * ( gelu(Agg_k) + P_k ) )  with Agg the masked per-hop sum of (x_j + edge_emb).
    With the geometric combine the whole expression inside mlp() is ONE kernel; [N,K,H] is never written.
    """

    def __init__(self, input_size, output_size, K, num_hop1_edge=1, num_pe=1, combine="independent"):
        super(KPGINPlusConv, self).__init__()
        self.aggr = "add"
        self.K = K
        self.output_size = output_size
        self.mlp = nn.Sequential(   # Linear-BN-ReLU x2 (KPGINplus.py:25-30); ReLUs folded into the BN kernels
            SplitKLinear(input_size, output_size), FusedBatchNorm1d(output_size, relu=True), nn.Identity(),
            SplitKLinear(output_size, output_size), FusedBatchNorm1d(output_size, relu=True), nn.Identity())
        self.hop1_edge_emb = torch.nn.Embedding(num_hop1_edge + 2, input_size, padding_idx=0)
        if self.K > 1:
            self.hopk_edge_emb = torch.nn.Embedding(num_pe + 2, input_size, padding_idx=0)
            self.hopk_node_path_emb = torch.nn.Embedding(num_pe, input_size, padding_idx=0)
            self.combine = make_combine(combine, self.K, self.output_size)
        else:
            self.hopk_edge_emb = None
            self.combine = torch.squeeze
        self.reset_parameters()

    def reset_parameters(self):
        self.hop1_edge_emb.reset_parameters()
        self.mlp.apply(self.weights_init)
        if self.K > 1:
            self.hopk_edge_emb.reset_parameters()
            self.hopk_node_path_emb.reset_parameters()
            self.combine.reset_parameters()

    def weights_init(self, m):
        if hasattr(m, "reset_parameters"):
            m.reset_parameters()

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None, post_norm=None, residual=None):
        """Reference signature (KPGINplus.py:61) plus two optional keywords used by this repo's backbone:
        `post_norm` (the BatchNorm GNNs.py:430 applies to the layer output) and `residual` (GNNs.py:436) are folded
        into the layer's dense-block kernel; returns post_norm(mlp(.)) + residual."""
        self._check_hops(edge_attr)
        plan, k = get_plan(edge_index, edge_attr, x.size(0))
        x = self._add_path_encoding(x, pe_attr)
        t0, tk = self._tables()
        if isinstance(self.combine, GeometricCombine):
            h = khop_aggregate(x, plan, k, P=peripheral_attr, T0=t0, Tk=tk, theta=self.combine.thetas(),
                               act=ACT_GELU, fuse=True)
        else:
            h = self.combine(khop_aggregate(x, plan, k, P=peripheral_attr, T0=t0, Tk=tk, act=ACT_GELU))
        bn3 = getattr(post_norm, "module", post_norm)          # backbone wraps its norms (state_dict key `module.*`)
        out = fused_dense_block(h, self.mlp[0], self.mlp[1], self.mlp[3], self.mlp[4], bn3, residual)
        if out is not None:
            return out
        out = self.mlp(h)
        if post_norm is not None:
            out = post_norm(out)
        if residual is not None:
            out = out + residual
        return out
