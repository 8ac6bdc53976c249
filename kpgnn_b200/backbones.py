"""Callers of the K-hop path for the BASELINE.json configs other than the headline one: the reference's `GNN`
(models/GNNs.py:22-236: KP-GIN / KP-GCN / KP-GraphSAGE, the same K-hop layer at every depth) and `GNNPrime`
(models/GNNs.py:478-723: K-hop layer(s) first, GINE layers after) backbones plus its graph-level heads
(models/GraphRegression.py:10-51, GraphClassification.py:10-52), restated from scratch with parameter names equal to
the reference's state_dict keys, so reference checkpoints (tests/golden/models_cfg.npz) load unchanged.

Like kpgnn_b200/model.py these exist so that bench.py, smoke() and the golden tests have a training step on the GPU
box, where the reference tree is absent; the reference's own unmodified models/GNNs.py runs on the same drop-in layers
(tests/test_reference_models_gpu.py).  Supported: norm_type Batch, virtual_node False, use_rd False,
JK in {last, concat, sum} -- what the reference's train scripts use for these configs; anything else raises.
"""
import copy

import torch
import torch.nn as nn

from .layers.feature_encoder import FeatureConcatEncoder
from .layers.gine import GINEConv
from .layers.norm import FusedBatchNorm1d
from .model import _Norm, segment_sum


def _clones(module, n):
    return nn.ModuleList(copy.deepcopy(module) for _ in range(n))


class _KHopBackboneBase(nn.Module):
    def _setup(self, num_layer, gnn_layer, init_emb, num_hop1_edge, max_edge_count, max_hop_num, max_distance_count,
               JK, norm_type, virtual_node, residual, use_rd, drop_prob):
        if norm_type != "Batch" or virtual_node or use_rd or JK not in ("last", "concat", "sum"):
            raise ValueError("kpgnn_b200.backbones supports norm_type=Batch, virtual_node=False, use_rd=False, "
                             "JK in (last, concat, sum); run the reference's models/GNNs.py over the drop-in layers "
                             "for other settings")
        self.num_layer, self.hidden_size, self.K = num_layer, gnn_layer.output_size, gnn_layer.K
        self.output_dk = gnn_layer.output_dk
        self.JK, self.residual = JK, residual
        self.dropout = nn.Dropout(drop_prob)
        width = (num_layer + 1) * self.hidden_size if JK == "concat" else self.hidden_size
        self.output_proj = nn.Sequential(nn.Linear(width, self.hidden_size), nn.ReLU(), nn.Dropout(drop_prob))
        self.init_proj = init_emb
        self.peripheral_edge_embedding = FeatureConcatEncoder([num_hop1_edge + 2, max_edge_count + 1], self.output_dk,
                                                              padding=0)
        self.pew = nn.Parameter(torch.rand(1))
        self.peripheral_configuration_embedding = FeatureConcatEncoder(
            [max_distance_count + 1 for _ in range(max_hop_num + 1)], self.output_dk, padding=0)
        self.pcw = nn.Parameter(torch.rand(1))

    def _reset_common(self):
        self.init_proj.reset_parameters()
        for m in self.output_proj:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()
        self.peripheral_edge_embedding.reset_parameters()
        self.peripheral_configuration_embedding.reset_parameters()
        nn.init.normal_(self.pew)
        nn.init.normal_(self.pcw)
        for n in self.norms:
            n.reset_parameters()

    def peripheral(self, data, num_nodes, like):
        """GNNs.py:172-179: sigmoid gates (GNNPlus uses tanh), encoders of width output_dk.  Both attribute sets and a
        kernel-eligible width: one fused gather-sum (kpgnn_b200/encoders.py); otherwise the reference's op-by-op form."""
        pea = getattr(data, "peripheral_edge_attr", None)
        pca = getattr(data, "peripheral_configuration_attr", None)
        dk = self.output_dk
        if pea is not None and pca is not None and like.is_cuda and dk % 4 == 0 and 4 <= dk <= 128:
            from .encoders import fused_peripheral_attr, peripheral_index
            return fused_peripheral_attr(self.peripheral_edge_embedding, self.peripheral_configuration_embedding,
                                         self.pew, self.pcw, peripheral_index(pea, pca), num_nodes, self.K,
                                         pea.size(2), gate="sigmoid")
        P = torch.zeros((num_nodes, self.K, dk), device=like.device, dtype=like.dtype)
        if pea is not None:
            P = P + torch.sigmoid(self.pew) * self.peripheral_edge_embedding(pea).sum(-2)
        if pca is not None:
            P = P + torch.sigmoid(self.pcw) * self.peripheral_configuration_embedding(pca)
        return P

    def _jk(self, h_list):
        if self.JK == "concat":
            rep = torch.cat(h_list, dim=1)
        elif self.JK == "last":
            rep = h_list[-1]
        else:
            rep = torch.stack(h_list, 0).sum(0)
        return self.output_proj(rep)


class KPGNNBackbone(_KHopBackboneBase):
    """GNN (models/GNNs.py:22-236)."""

    def __init__(self, num_layer, gnn_layer, init_emb, num_hop1_edge, max_edge_count, max_hop_num, max_distance_count,
                 JK="last", norm_type="Batch", virtual_node=False, residual=False, use_rd=False, drop_prob=0.0):
        super().__init__()
        self._setup(num_layer, gnn_layer, init_emb, num_hop1_edge, max_edge_count, max_hop_num, max_distance_count,
                    JK, norm_type, virtual_node, residual, use_rd, drop_prob)
        self.gnns = _clones(gnn_layer, num_layer)
        self.norms = nn.ModuleList(_Norm(self.hidden_size) for _ in range(num_layer))
        self.reset_parameters()

    def reset_parameters(self):
        self._reset_common()
        for g in self.gnns:
            g.reset_parameters()

    def forward(self, data):
        x = self.init_proj(data).squeeze()
        P = self.peripheral(data, x.size(0), x)
        pe = getattr(data, "pe_attr", None)
        h_list = [x]
        for l in range(self.num_layer):
            h = self.gnns[l](h_list[l], data.edge_index, data.edge_attr, pe, P)
            h = self.norms[l](h)
            if l != self.num_layer - 1:
                h = self.dropout(h)
            if self.residual:
                h = h + h_list[l]
            h_list.append(h)
        return self._jk(h_list)


class KPGNNPrimeBackbone(_KHopBackboneBase):
    """GNNPrime (models/GNNs.py:478-723): num_l1_layer K-hop layers, then GINE layers on the hop-1 column."""

    def __init__(self, num_layer, gnn_layer, init_emb, num_hop1_edge, max_edge_count, max_hop_num, max_distance_count,
                 num_l1_layer=1, JK="last", norm_type="Batch", virtual_node=False, residual=False, use_rd=False,
                 drop_prob=0.0):
        super().__init__()
        assert num_l1_layer > 0 and num_layer >= 2
        self.num_l1_layer, self.num_l2_layer = num_l1_layer, num_layer - num_l1_layer
        self._setup(num_layer, gnn_layer, init_emb, num_hop1_edge, max_edge_count, max_hop_num, max_distance_count,
                    JK, norm_type, virtual_node, residual, use_rd, drop_prob)
        self.khop_gnns = _clones(gnn_layer, num_l1_layer)
        self.gins = _clones(GINEConv(self.hidden_size, self.hidden_size, num_hop1_edge=num_hop1_edge),
                            self.num_l2_layer)
        self.norms = nn.ModuleList(_Norm(self.hidden_size) for _ in range(num_layer))
        self.reset_parameters()

    def reset_parameters(self):
        self._reset_common()
        for g in list(self.khop_gnns) + list(self.gins):
            g.reset_parameters()

    def forward(self, data):
        x = self.init_proj(data).squeeze()
        P = self.peripheral(data, x.size(0), x)
        pe = getattr(data, "pe_attr", None)
        h_list = [x]
        for l in range(self.num_layer):
            if l < self.num_l1_layer:
                h = self.khop_gnns[l](h_list[l], data.edge_index, data.edge_attr, pe, P)
            else:
                h = self.gins[l - self.num_l1_layer](h_list[l], data.edge_index, data.edge_attr[:, :1])
            h = self.norms[l](h)
            # GNNs.py:657 applies dropout after EVERY K-hop layer, :681 after every GINE layer but the last
            if l < self.num_l1_layer or l != self.num_layer - 1:
                h = self.dropout(h)
            if self.residual:
                h = h + h_list[l]
            h_list.append(h)
        return self._jk(h_list)


class GraphHead(nn.Module):
    """GraphRegression (output_size None -> `regressor`, squeezed) / GraphClassification (`classifier`), sum or mean
    pooling through the ordered segment-sum kernel."""

    def __init__(self, embedding_model, pooling_method="sum", output_size=None):
        super().__init__()
        if pooling_method not in ("sum", "mean"):
            raise ValueError("The pooling method not implemented")
        self.embedding_model, self.pooling_method = embedding_model, pooling_method
        if output_size is None:
            self.regressor = nn.Linear(embedding_model.hidden_size, 1)
        else:
            self.classifier = nn.Linear(embedding_model.hidden_size, output_size)
        self.output_size = output_size

    def forward(self, data):
        h = self.embedding_model(data)
        ng = data.num_graphs if getattr(data, "num_graphs", None) is not None else int(data.batch[-1]) + 1
        pooled = segment_sum(h, data.batch, ng, mean=self.pooling_method == "mean")
        return self.regressor(pooled).squeeze() if self.output_size is None else self.classifier(pooled)


def make_model(model_name, hidden_size, K, num_layer, input_size, num_hop1_edge, max_pe_num, max_edge_count, max_hop_num,
               max_distance_count, combine="geometric", JK="last", residual=False, output_size=None,
               pooling_method="sum", drop_prob=0.0, eps=0.0, train_eps=False, aggr="add"):
    """get_model() of the reference's train scripts (train_EXP.py / train_SR.py / train_ZINC.py) on the product."""
    import argparse
    from .layers.input_encoder import EmbeddingEncoder
    from .layers.layer_utils import make_gnn_layer
    from .model import KPGNNPlusBackbone
    if model_name == "KPGINPlus":
        gnn = KPGNNPlusBackbone(num_layer, hidden_size, K, input_size, num_hop1_edge, max_pe_num, max_edge_count,
                                max_hop_num, max_distance_count, combine, JK, residual, drop_prob)
    else:
        args = argparse.Namespace(model_name=model_name, hidden_size=hidden_size, K=K, num_hop1_edge=num_hop1_edge,
                                  max_pe_num=max_pe_num, combine=combine, num_layer=num_layer, eps=eps,
                                  train_eps=train_eps, aggr=aggr)
        cls = KPGNNPrimeBackbone if model_name == "KPGINPrime" else KPGNNBackbone
        gnn = cls(num_layer, make_gnn_layer(args), EmbeddingEncoder(input_size, hidden_size), num_hop1_edge,
                  max_edge_count, max_hop_num, max_distance_count, JK=JK, residual=residual, drop_prob=drop_prob)
    return GraphHead(gnn, pooling_method, output_size)
