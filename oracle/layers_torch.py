"""ORACLE (test infrastructure, not product code): plain PyTorch fp32 restatement of the reference's K-hop
aggregation layers, with the reference's dense `[E_K, k, d]` message tensors, run on the CPU.

Follows /root/reference/layers/{KPGIN.py:12-121, KPGINplus.py:10-88, KPGCN.py:11-126, KPGraphSAGE.py:12-106,
gine.py:9-59, combine.py:8-58} and the PyG `MessagePassing.propagate` contract they call (third-party,
torch_geometric pinned 2.1.0, README.md:12, not vendored): flow source->target, `x_j = x[edge_index[0]]`,
sum (or mean) at `edge_index[1]`, `dim_size = N`.

Parity status: PINNED against the reference itself (imported unmodified behind oracle/pyg_standin in the build
container): tests/test_oracle_cpu.py compares live when /root/reference exists, and tests/golden/layers.npz
hold the reference's own outputs/gradients (made by oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
Parameter names equal the reference's state_dict keys so one state_dict loads into reference, oracle and product.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------
# message + aggregate, the part every layer shares  (KPGIN.py:115-118 and siblings)
# --------------------------------------------------------------------------------------------------------------
def dense_khop_aggregate(x, edge_index, edge_attr, hop1_table, hopk_table, norm=None, aggr="add"):
    """x [N,k,d]; edge_attr [E,k] long; returns [N,k,d].  Materialises the [E,k,d] message tensor exactly
    like the reference does."""
    src, dst = edge_index[0], edge_index[1]
    msg = x.index_select(0, src)
    if hop1_table is not None:
        emb = F.embedding(edge_attr[:, :1], hop1_table, padding_idx=0)
        if edge_attr.size(1) > 1:
            emb = torch.cat([emb, F.embedding(edge_attr[:, 1:], hopk_table, padding_idx=0)], dim=-2)
        msg = msg + emb
    if norm is not None:
        msg = norm.unsqueeze(-1) * msg
    msg = msg.masked_fill(edge_attr.unsqueeze(-1) == 0, 0.)
    out = torch.zeros((x.size(0),) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device).index_add_(0, dst, msg)
    if aggr == "mean":
        cnt = torch.zeros(x.size(0), dtype=msg.dtype, device=msg.device).index_add_(
            0, dst, torch.ones(dst.numel(), dtype=msg.dtype, device=msg.device))
        out = out / cnt.clamp_(min=1).view(-1, 1, 1)
    elif aggr not in ("add", "sum"):
        raise NotImplementedError(aggr)
    return out


# --------------------------------------------------------------------------------------------------------------
# combine.py
# --------------------------------------------------------------------------------------------------------------
class OracleGeometricCombine(nn.Module):
    """combine.py:30-58"""

    def __init__(self, K, hidden_size):
        super().__init__()
        self.alphas = nn.Parameter(torch.zeros(hidden_size))
        self.K = K

    def reset_parameters(self):
        nn.init.zeros_(self.alphas)

    def thetas(self):
        a = torch.sigmoid(self.alphas)
        powers = torch.arange(self.K, dtype=a.dtype, device=a.device).view(-1, 1)
        return torch.softmax(a.unsqueeze(0) * (1 - a).unsqueeze(0) ** powers, dim=0)      # [K, d]

    def forward(self, x):
        return (x * self.thetas().unsqueeze(0)).sum(dim=-2)


class OracleAttentionCombine(nn.Module):
    """combine.py:8-27"""

    def __init__(self, hidden_size, K):
        super().__init__()
        self.attention_lstm = nn.LSTM(hidden_size, K, 1, batch_first=True, bidirectional=True, dropout=0.)

    def reset_parameters(self):
        self.attention_lstm.reset_parameters()

    def forward(self, x):
        score, _ = self.attention_lstm(x)
        w = torch.softmax(score.sum(-1), dim=1).unsqueeze(-1)
        return (x * w).sum(1)


def _make_combine(kind, K, d):
    if kind == "attention":
        return OracleAttentionCombine(d, K)
    if kind == "geometric":
        return OracleGeometricCombine(K, d)
    raise ValueError("Not implemented combine function")


def _squeeze(x):
    return torch.squeeze(x)


class _KHopBase(nn.Module):
    def _tables(self):
        return self.hop1_edge_emb.weight, (self.hopk_edge_emb.weight if self.hopk_edge_emb is not None else None)

    def _add_path_encoding(self, x, pe_attr):
        # reference does this in place on a view of the caller's tensor (KPGIN.py:92-94)
        if self.K > 1 and pe_attr is not None:
            x[:, 1:] = x[:, 1:] + self.hopk_node_path_emb(pe_attr)
        return x


class OracleKPGINConv(_KHopBase):
    """KPGIN.py:12-121"""

    def __init__(self, input_size, output_size, K, eps=0., train_eps=False, num_hop1_edge=1, num_pe=1,
                 combine="geometric"):
        super().__init__()
        assert input_size % K == 0 and output_size % K == 0
        self.K, self.output_size = K, output_size
        self.input_dk, self.output_dk = input_size // K, output_size // K
        self.hop_proj1 = nn.Parameter(torch.empty(K, self.input_dk, self.output_dk))
        self.hop_bias1 = nn.Parameter(torch.empty(K, self.output_dk))
        self.hop_proj2 = nn.Parameter(torch.empty(K, self.output_dk, self.output_dk))
        self.hop_bias2 = nn.Parameter(torch.empty(K, self.output_dk))
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))
        self.hop1_edge_emb = nn.Embedding(num_hop1_edge + 2, self.input_dk, padding_idx=0)
        if K > 1:
            self.hopk_edge_emb = nn.Embedding(num_pe + 2, self.input_dk, padding_idx=0)
            self.hopk_node_path_emb = nn.Embedding(num_pe, self.input_dk, padding_idx=0)
            self.combine_proj = nn.Linear(self.output_dk, output_size)
            self.combine = _make_combine(combine, K, self.output_dk)
        else:
            self.hopk_edge_emb = None
            self.combine = _squeeze
            self.combine_proj = nn.Identity()
        for p in (self.hop_proj1, self.hop_proj2):
            nn.init.kaiming_uniform_(p)
        for w, b in ((self.hop_proj1, self.hop_bias1), (self.hop_proj2, self.hop_bias2)):
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(w)
            nn.init.uniform_(b, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        x = self._add_path_encoding(x.view(-1, self.K, self.input_dk), pe_attr)
        t1, tk = self._tables()
        z = dense_khop_aggregate(x, edge_index, edge_attr, t1, tk)
        if peripheral_attr is not None:
            z = z + peripheral_attr
        z = (z + (1 + self.eps) * x).permute(1, 0, 2)
        z = F.relu(torch.matmul(z, self.hop_proj1) + self.hop_bias1.unsqueeze(1))
        z = F.relu(torch.matmul(z, self.hop_proj2) + self.hop_bias2.unsqueeze(1))
        return self.combine_proj(self.combine(z.permute(1, 0, 2)))


def _mlp(i, o):
    return nn.Sequential(nn.Linear(i, o), nn.BatchNorm1d(o), nn.ReLU(), nn.Linear(o, o), nn.BatchNorm1d(o), nn.ReLU())


class OracleKPGINPlusConv(_KHopBase):
    """KPGINplus.py:10-88"""

    def __init__(self, input_size, output_size, K, num_hop1_edge=1, num_pe=1, combine="independent"):
        super().__init__()
        self.K, self.output_size = K, output_size
        self.mlp = _mlp(input_size, output_size)
        self.hop1_edge_emb = nn.Embedding(num_hop1_edge + 2, input_size, padding_idx=0)
        if K > 1:
            self.hopk_edge_emb = nn.Embedding(num_pe + 2, input_size, padding_idx=0)
            self.hopk_node_path_emb = nn.Embedding(num_pe, input_size, padding_idx=0)
            self.combine = _make_combine(combine, K, output_size)
        else:
            self.hopk_edge_emb = None
            self.combine = _squeeze

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        x = self._add_path_encoding(x, pe_attr)
        t1, tk = self._tables()
        z = F.gelu(dense_khop_aggregate(x, edge_index, edge_attr, t1, tk))
        if peripheral_attr is not None:
            z = z + peripheral_attr
        return self.mlp(self.combine(z))


class OracleKPGCNConv(_KHopBase):
    """KPGCN.py:11-126"""

    def __init__(self, input_size, output_size, K, num_hop1_edge=1, num_pe=1, combine="geometric"):
        super().__init__()
        assert output_size % K == 0
        self.K, self.output_size, self.output_dk = K, output_size, output_size // K
        self.hop_proj = nn.Linear(input_size, output_size)
        self.hop1_edge_emb = nn.Embedding(num_hop1_edge + 2, self.output_dk, padding_idx=0)
        if K > 1:
            self.hopk_edge_emb = nn.Embedding(num_pe + 2, self.output_dk, padding_idx=0)
            self.hopk_node_path_emb = nn.Embedding(num_pe, self.output_dk, padding_idx=0)
            self.combine_proj = nn.Linear(self.output_dk, output_size)
            self.combine = _make_combine(combine, K, self.output_dk)
        else:
            self.hopk_edge_emb = None
            self.combine = _squeeze
            self.combine_proj = nn.Identity()

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        n = x.size(0)
        loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).unsqueeze(0).repeat(2, 1)
        edge_index = torch.cat([edge_index, loops], dim=1)                                  # :85
        edge_attr = torch.cat([edge_attr, torch.ones(n, self.K, dtype=edge_attr.dtype, device=edge_attr.device)], 0)  # :87-89
        x = self._add_path_encoding(self.hop_proj(x).view(-1, self.K, self.output_dk), pe_attr)
        src, dst = edge_index
        deg = torch.zeros(n, self.K, device=x.device).index_add_(0, dst, (edge_attr > 0).float())   # :11-25
        dis = deg.pow(-0.5)
        norm = dis[src] * dis[dst]
        t1, tk = self._tables()
        z = F.relu(dense_khop_aggregate(x, edge_index, edge_attr, t1, tk, norm=norm))
        if peripheral_attr is not None:
            z = z + peripheral_attr
        return self.combine_proj(self.combine(z))


class OracleKPGraphSAGEConv(_KHopBase):
    """KPGraphSAGE.py:12-106"""

    def __init__(self, input_size, output_size, K, aggr="mean", num_hop1_edge=1, num_pe=1, combine="geometric"):
        super().__init__()
        assert input_size % K == 0 and output_size % K == 0
        self.aggr, self.K, self.output_size = aggr, K, output_size
        self.input_dk, self.output_dk = input_size // K, output_size // K
        self.hop_proj = nn.Parameter(torch.empty(K, 2 * self.input_dk, self.output_dk))
        self.hop_bias = nn.Parameter(torch.empty(K, self.output_dk))
        self.hop1_edge_emb = nn.Embedding(num_hop1_edge + 2, self.input_dk, padding_idx=0)
        if K > 1:
            self.combine_proj = nn.Linear(self.output_dk, output_size)
            self.hopk_edge_emb = nn.Embedding(num_pe + 2, self.input_dk, padding_idx=0)
            self.hopk_node_path_emb = nn.Embedding(num_pe, self.input_dk, padding_idx=0)
            self.combine = _make_combine(combine, K, self.output_dk)
        else:
            self.hopk_edge_emb = None
            self.combine = _squeeze
            self.combine_proj = nn.Identity()
        nn.init.kaiming_uniform_(self.hop_proj)
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.hop_proj)
        nn.init.uniform_(self.hop_bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def forward(self, x, edge_index, edge_attr, pe_attr=None, peripheral_attr=None):
        x = self._add_path_encoding(x.view(-1, self.K, self.input_dk), pe_attr)
        t1, tk = self._tables()
        z = dense_khop_aggregate(x, edge_index, edge_attr, t1, tk, aggr=self.aggr)
        if peripheral_attr is not None:
            z = z + peripheral_attr
        y = torch.cat([x, z], dim=-1).permute(1, 0, 2)
        y = (torch.matmul(y, self.hop_proj) + self.hop_bias.unsqueeze(1)).permute(1, 0, 2)
        y = F.normalize(F.relu(y), p=2, dim=-1)
        return self.combine_proj(self.combine(y))


class OracleGINEConv(nn.Module):
    """gine.py:9-59"""

    def __init__(self, input_size, output_size, eps=0., num_hop1_edge=1, train_eps=False):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))
        self.mlp = _mlp(input_size, output_size)
        self.hop1_edge_emb = nn.Embedding(num_hop1_edge + 2, input_size, padding_idx=0)

    def forward(self, x, edge_index, edge_attr):
        x = x.view(-1, 1, self.input_size)
        out = dense_khop_aggregate(x, edge_index, edge_attr, self.hop1_edge_emb.weight, None)
        return self.mlp((out + (1 + self.eps) * x).squeeze())


class OracleKGINConv(nn.Module):
    """run_simulation.py:29-93 (the script itself cannot be imported here: it needs matplotlib; this class is
    checked against the reference by reading only -- parity UNPINNED for this one module)."""

    def __init__(self, hidden_size, K, eps=0., train_eps=False):
        super().__init__()
        self.K, self.hidden_size = K, hidden_size
        self.proj = nn.Linear(1, K * hidden_size)
        self.hop_proj1 = nn.Parameter(torch.empty(K, hidden_size, hidden_size))
        self.hop_bias1 = nn.Parameter(torch.empty(K, hidden_size))
        self.hop_proj2 = nn.Parameter(torch.empty(K, hidden_size, hidden_size))
        self.hop_bias2 = nn.Parameter(torch.empty(K, hidden_size))
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))
        self.combine_proj = nn.Linear(hidden_size * K, hidden_size)

    def forward(self, x, edge_index, edge_attr, batch):
        x = self.proj(x).view(-1, self.K, self.hidden_size)
        z = dense_khop_aggregate(x, edge_index, edge_attr, None, None) + (1 + self.eps) * x   # :73-75
        z = z.permute(1, 0, 2)
        z = F.relu(torch.matmul(z, self.hop_proj1) + self.hop_bias1.unsqueeze(1))
        z = F.relu(torch.matmul(z, self.hop_proj2) + self.hop_bias2.unsqueeze(1))
        return self.combine_proj(z.permute(1, 0, 2).contiguous().view(-1, self.K * self.hidden_size))


def make_oracle_layer(model_name, hidden, K, num_layer=None, eps=0., train_eps=False, num_hop1_edge=1, max_pe_num=1,
                      combine="geometric", aggr="add"):
    """layer_utils.py:10-34"""
    if model_name == "KPGCN":
        return OracleKPGCNConv(hidden, hidden, K, num_hop1_edge, max_pe_num, combine)
    if model_name in ("KPGIN", "KPGINPrime"):
        return OracleKPGINConv(hidden, hidden, K, eps, train_eps, num_hop1_edge, max_pe_num, combine)
    if model_name == "KPGraphSAGE":
        return OracleKPGraphSAGEConv(hidden, hidden, K, aggr, num_hop1_edge, max_pe_num, combine)
    if model_name == "KPGINPlus":
        return [OracleKPGINPlusConv(hidden, hidden, min(l, K), num_hop1_edge, max_pe_num, combine)
                for l in range(1, num_layer + 1)]
    raise ValueError("Not supported GNN type")
