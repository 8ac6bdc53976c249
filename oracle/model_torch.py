"""ORACLE (test infrastructure, not product code): plain PyTorch restatement of the reference's KP-GIN+ graph
regression model, `GraphRegression(GNNPlus(...))` -- /root/reference/models/GNNs.py:238-474 (norm_type="Batch",
virtual_node=False, use_rd=False) and models/GraphRegression.py:10-51 with sum pooling -- on the oracle layers of
oracle/layers_torch.py (dense [E,k,d] message tensors, as the reference computes them).

Used as (a) the checker for kpgnn_b200/model.py, (b) the CPU baseline bench.py times (`cpu_baseline`,
`--impl reference`).  Parity status: PINNED -- tests/golden/model_zinc.npz holds the reference model's own
output, loss and gradients for a committed state_dict and batch (oracle/make_golden.py); state_dict keys equal
the reference's.  Never imported by kpgnn_b200/.
"""
import torch
import torch.nn as nn

from .layers_torch import OracleKPGINPlusConv


class _ConcatEncoder(nn.Module):
    """layers/feature_encoder.py:37-67.  GNNs.py:312,316 pass `padding=0`, which is falsy, so the reference's
    embeddings have NO padding row: row 0 is trainable and receives gradient."""

    def __init__(self, dims, hidden):
        super().__init__()
        self.embedding_list = nn.ModuleList(nn.Embedding(d, hidden) for d in dims)
        self.proj = nn.Linear(len(dims) * hidden, hidden)

    def forward(self, x):
        return self.proj(torch.cat([e(x[..., i]) for i, e in enumerate(self.embedding_list)], dim=-1))


class _Emb(nn.Module):
    """layers/input_encoder.py:9-23"""

    def __init__(self, n, hidden):
        super().__init__()
        self.init_proj = nn.Embedding(n, hidden)

    def forward(self, x):
        return self.init_proj(x)


class _BN(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.module = nn.BatchNorm1d(w)

    def forward(self, x):
        return self.module(x)


class OracleGNNPlus(nn.Module):
    def __init__(self, num_layer, hidden_size, K, input_size, num_hop1_edge, max_pe_num, max_edge_count,
                 max_hop_num, max_distance_count, combine="geometric", JK="concat", residual=True):
        super().__init__()
        self.num_layer, self.hidden_size, self.K, self.JK, self.residual = num_layer, hidden_size, K, JK, residual
        width = (num_layer + 1) * hidden_size if JK == "concat" else hidden_size
        self.output_proj = nn.Sequential(nn.Linear(width, hidden_size), nn.ReLU(), nn.Dropout(0.0))
        self.init_proj = _Emb(input_size, hidden_size)
        self.peripheral_edge_embedding = _ConcatEncoder([num_hop1_edge + 2, max_edge_count + 1], hidden_size)
        self.pew = nn.Parameter(torch.rand(1))
        self.peripheral_configuration_embedding = _ConcatEncoder([max_distance_count + 1] * (max_hop_num + 1),
                                                                 hidden_size)
        self.pcw = nn.Parameter(torch.rand(1))
        self.gnns = nn.ModuleList(OracleKPGINPlusConv(hidden_size, hidden_size, min(l, K), num_hop1_edge, max_pe_num,
                                                      combine) for l in range(1, num_layer + 1))
        self.norms = nn.ModuleList(_BN(hidden_size) for _ in range(num_layer))

    def forward(self, d):
        x = self.init_proj(d["x"]).squeeze()                                               # GNNs.py:385
        P = torch.zeros(x.size(0), self.K, self.hidden_size, dtype=x.dtype, device=x.device)
        if d.get("peripheral_edge_attr") is not None:                                      # :394-396
            P = P + torch.tanh(self.pew) * self.peripheral_edge_embedding(d["peripheral_edge_attr"]).sum(-2)
        if d.get("peripheral_configuration_attr") is not None:                             # :398-400
            P = P + torch.tanh(self.pcw) * self.peripheral_configuration_embedding(d["peripheral_configuration_attr"])
        hs, last = [x], x
        for l in range(self.num_layer):                                                    # :410-438
            k = min(l + 1, self.K)
            xs = torch.cat([hs[j].unsqueeze(1) for j in range(l, l - k, -1)], dim=1)
            pe = d["pe_attr"][:, :k - 1] if d.get("pe_attr") is not None else None
            h = self.norms[l](self.gnns[l](xs, d["edge_index"], d["edge_attr"][:, :k], pe, P[:, :k]))
            if self.residual:
                h = h + last
                last = h
            hs.append(h)
        rep = torch.cat(hs, dim=1) if self.JK == "concat" else hs[-1]                      # :455-458
        return self.output_proj(rep)


class OracleGraphRegression(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.embedding_model = OracleGNNPlus(**kw)
        self.regressor = nn.Linear(self.embedding_model.hidden_size, 1)

    def forward(self, d):
        h = self.embedding_model(d)
        pooled = torch.zeros(d["num_graphs"], h.size(1), dtype=h.dtype, device=h.device).index_add_(0, d["batch"], h)
        return self.regressor(pooled).squeeze()


def zinc_oracle_model(K=8, num_layer=8, hidden=104, combine="geometric"):
    return OracleGraphRegression(num_layer=num_layer, hidden_size=hidden, K=K, input_size=21, num_hop1_edge=3,
                                 max_pe_num=50, max_edge_count=50, max_hop_num=6, max_distance_count=50,
                                 combine=combine, JK="concat", residual=True)


def l1_loss(score, y):
    return (score.squeeze() - y.squeeze()).abs().mean()                                    # train_ZINC.py:42
