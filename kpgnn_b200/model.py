"""Host-side backbone that drives the hot path for the headline workload:
KP-GIN+ graph regression = the reference's `GraphRegression(GNNPlus(...))` (models/GNNs.py:238-474,
models/GraphRegression.py:10-51), restated as one module whose parameter names equal the reference's
state_dict keys, so a reference checkpoint (or the golden state_dict under tests/golden/) loads unchanged.

This is the CALLER of the path, written from scratch so that bench.py / smoke() have a training step on the GPU
box where the reference tree does not exist; the reference's own `models/GNNs.py` runs on the same drop-in
layers unchanged.  Dense pieces (Embedding, Linear, BatchNorm) are library GEMMs on purpose.
"""
import torch
import torch.nn as nn

from .layers.KPGINplus import KPGINPlusConv
from .layers.KPGIN import KPGINConv
from .layers.gine import GINEConv
from .layers.feature_encoder import FeatureConcatEncoder
from .layers.input_encoder import EmbeddingEncoder
from .layers.norm import FusedBatchNorm1d
from .layers._base import SplitKLinear


class Batch(object):
    """Collated batch in the reference's wire layout (PyG `Batch.from_data_list`): plain attribute bag."""
    FIELDS = ("x", "edge_index", "edge_attr", "pe_attr", "peripheral_edge_attr", "peripheral_configuration_attr",
              "batch", "y")

    def __init__(self, **kw):
        self.num_graphs = kw.pop("num_graphs")
        self.num_nodes = kw.pop("num_nodes", None)
        self.n_dev = kw.pop("n_dev", None)     # padded-capacity batches: device int32 scalar with the true node count
        for f in self.FIELDS:
            setattr(self, f, kw.get(f))

    def __contains__(self, key):
        return getattr(self, key, None) is not None

    def to(self, device, non_blocking=False):
        out = Batch(num_graphs=self.num_graphs, num_nodes=self.num_nodes)
        for f in self.FIELDS:
            v = getattr(self, f)
            setattr(out, f, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def pin_memory(self):
        out = Batch(num_graphs=self.num_graphs, num_nodes=self.num_nodes)
        for f in self.FIELDS:
            v = getattr(self, f)
            setattr(out, f, v.pin_memory() if torch.is_tensor(v) else v)
        return out

    def nbytes(self):
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.FIELDS
                   if torch.is_tensor(getattr(self, f)))


class _Norm(nn.Module):
    """state_dict-compatible with torch_geometric.nn.BatchNorm (`module.*` keys), GNNs.py:324-325."""

    def __init__(self, width):
        super().__init__()
        self.module = FusedBatchNorm1d(width)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x):
        return self.module(x)


class _SegmentSum(torch.autograd.Function):
    """kp_segment_sum (include/kpgnn.h): ordered per-graph row sums over the sorted `batch` vector."""

    @staticmethod
    def forward(ctx, x, batch, num_graphs, mean):
        import ctypes as C
        from . import _lib
        lib = _lib.lib()
        if not x.is_cuda:
            raise _lib.KpError("kpgnn_b200 runs on CUDA tensors only (no CPU fallback)")
        x = x.detach()
        if x.stride(-1) != 1:
            x = x.contiguous()
        batch = batch.contiguous()
        out = torch.empty((num_graphs, x.size(1)), dtype=torch.float32, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.kp_segment_sum(x.data_ptr(), x.stride(0), batch.data_ptr(), x.size(0), x.size(1), num_graphs,
                                      1 if mean else 0, out.data_ptr(), st), "kp_segment_sum")
        ctx.batch, ctx.mean, ctx.num_graphs = batch, mean, num_graphs
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.mean:
            cnt = torch.bincount(ctx.batch, minlength=ctx.num_graphs).clamp_(min=1).to(dout.dtype)
            dout = dout / cnt.unsqueeze(1)
        # rows of a padded capacity carry graph id == num_graphs (kpgnn_b200/wire.py): they read a zero row
        dout = torch.cat([dout, dout.new_zeros(1, dout.size(1))], 0)
        return dout.index_select(0, ctx.batch), None, None, None


def segment_sum(x, batch, num_graphs, mean=False):
    """global_add_pool / global_mean_pool (GraphRegression.py:26) over the sorted `batch` vector: rows added in
    ascending order per graph, no float atomics (torch's index_add_ is a float atomicAdd, so the readout -- and with
    it the loss and every gradient -- would differ in the last bits from run to run)."""
    return _SegmentSum.apply(x, batch, num_graphs, mean)


class KPGNNPlusBackbone(nn.Module):
    """GNNPlus (models/GNNs.py:238-474) for norm_type=Batch, virtual_node=False, use_rd=False."""

    def __init__(self, num_layer, hidden_size, K, input_size, num_hop1_edge, max_pe_num, max_edge_count,
                 max_hop_num, max_distance_count, combine="geometric", JK="concat", residual=True, drop_prob=0.0):
        super().__init__()
        assert num_layer >= K
        self.num_layer, self.hidden_size, self.K = num_layer, hidden_size, K
        self.JK, self.residual = JK, residual
        self.dropout = nn.Dropout(drop_prob)
        width = (num_layer + 1) * hidden_size if JK == "concat" else hidden_size
        self.output_proj = nn.Sequential(SplitKLinear(width, hidden_size), nn.ReLU(), nn.Dropout(drop_prob))
        self.init_proj = EmbeddingEncoder(input_size, hidden_size)
        self.peripheral_edge_embedding = FeatureConcatEncoder([num_hop1_edge + 2, max_edge_count + 1], hidden_size,
                                                              padding=0)
        self.pew = nn.Parameter(torch.rand(1))
        self.peripheral_configuration_embedding = FeatureConcatEncoder(
            [max_distance_count + 1 for _ in range(max_hop_num + 1)], hidden_size, padding=0)
        self.pcw = nn.Parameter(torch.rand(1))
        self.gnns = nn.ModuleList(KPGINPlusConv(hidden_size, hidden_size, min(l, K), num_hop1_edge, max_pe_num,
                                                combine) for l in range(1, num_layer + 1))
        self.norms = nn.ModuleList(_Norm(hidden_size) for _ in range(num_layer))
        self.reset_parameters()

    def reset_parameters(self):
        self.init_proj.reset_parameters()
        for m in self.output_proj:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()
        self.peripheral_edge_embedding.reset_parameters()
        self.peripheral_configuration_embedding.reset_parameters()
        nn.init.normal_(self.pew)
        nn.init.normal_(self.pcw)
        for g in self.gnns:
            g.reset_parameters()
        for n in self.norms:
            n.reset_parameters()

    def peripheral(self, data, num_nodes, like):
        """GNNs.py:393-400: integer peripheral attributes -> [N,K,H] floats (tanh gates in GNNPlus).
        Both attribute sets present (the normal case): one fused gather-sum kernel (kpgnn_b200/encoders.py);
        otherwise the reference's op-by-op form."""
        if data.peripheral_edge_attr is not None and data.peripheral_configuration_attr is not None:
            from .encoders import fused_peripheral_attr, peripheral_index
            idx = getattr(data, "_peripheral_idx", None)
            ver = (data.peripheral_edge_attr._version, data.peripheral_configuration_attr._version)
            if idx is None or idx[0] != ver:
                idx = (ver, peripheral_index(data.peripheral_edge_attr, data.peripheral_configuration_attr))
                data._peripheral_idx = idx
            return fused_peripheral_attr(self.peripheral_edge_embedding, self.peripheral_configuration_embedding,
                                         self.pew, self.pcw, idx[1], num_nodes, self.K,
                                         data.peripheral_edge_attr.size(2), gate="tanh")
        P = torch.zeros((num_nodes, self.K, self.hidden_size), device=like.device, dtype=like.dtype)
        if data.peripheral_edge_attr is not None:
            P = P + torch.tanh(self.pew) * self.peripheral_edge_embedding(data.peripheral_edge_attr).sum(-2)
        if data.peripheral_configuration_attr is not None:
            P = P + torch.tanh(self.pcw) * self.peripheral_configuration_embedding(data.peripheral_configuration_attr)
        return P

    def input_embedding(self, data):
        """init_proj(data) (GNNs.py:385).  An integer [N] input is a 1-slot gather-sum: same kernel as the
        peripheral stage, whose deterministic gradient replaces the sort-based embedding backward."""
        xin = data.x
        if xin.dim() == 1 and xin.dtype == torch.int64 and xin.is_cuda and self.hidden_size % 4 == 0 \
                and self.hidden_size <= 128:
            from .encoders import _TableSum
            w = self.init_proj.init_proj.weight
            return _TableSum.apply(w, xin.view(-1, 1), [0], [0, 1], [0, w.size(0)])
        return self.init_proj(data).squeeze()

    def _tail_fusable(self):
        """Is the tail Linear -> ReLU -> identity dropout (what the fused regression head can absorb)?"""
        mods = list(self.output_proj)
        return len(mods) == 3 and isinstance(mods[1], nn.ReLU) and isinstance(mods[2], nn.Dropout) and \
            (mods[2].p == 0.0 or not self.training)

    def _tail(self, rep_in, pre_activation):
        """output_proj on the JK representation; with pre_activation only its Linear (the fused head applies the ReLU)."""
        return self.output_proj[0](rep_in) if pre_activation else self.output_proj(rep_in)

    def forward(self, data, pre_activation=False):
        if pre_activation and not self._tail_fusable():
            return None
        x = self.input_embedding(data)
        N = x.size(0)
        P = self.peripheral(data, N, x)
        fused = self._forward_stack(data, x, P, pre_activation)
        if fused is not None:
            return fused
        h_list = [x]
        last_h = x
        for l in range(self.num_layer):
            k = min(l + 1, self.K)
            # newest layer first, GNNs.py:413-418
            xs = torch.stack([h_list[j] for j in range(l, l - k, -1)], dim=1)
            pe = data.pe_attr[:, :k - 1] if data.pe_attr is not None else None
            last = l == self.num_layer - 1
            if self.training and (last or self.dropout.p == 0.0):
                # norm + (identity dropout) + residual ride in the layer's dense-block kernel (GNNs.py:430-438)
                h = self.gnns[l](xs, data.edge_index, data.edge_attr[:, :k], pe, P[:, :k], post_norm=self.norms[l],
                                 residual=last_h if self.residual else None)
                if self.residual:
                    last_h = h
            else:
                h = self.gnns[l](xs, data.edge_index, data.edge_attr[:, :k], pe, P[:, :k])
                h = self.norms[l](h)
                if not last:
                    h = self.dropout(h)
                if self.residual:
                    h = h + last_h
                    last_h = h
            h_list.append(h)
        if self.JK == "concat":
            rep = torch.cat(h_list, dim=1)
        elif self.JK == "last":
            rep = h_list[-1]
        elif self.JK == "sum":
            rep = torch.stack(h_list, 0).sum(0)
        else:
            raise ValueError("JK=%r not supported by this backbone" % (self.JK,))
        return self._tail(rep, pre_activation)

    use_stack = True          # tests switch this off to compare against the layer-by-layer path

    def _forward_stack(self, data, x, P, pre_activation=False):
        """All layers as one autograd node over a layer-history buffer (kpgnn_b200/stack.py); None if not applicable."""
        from .stack import kpginplus_stack, stack_applicable
        from .layers._base import _all_zero, SplitKLinear  # noqa: F401
        from .layers._base import _SplitKLinearFn
        from .plan import get_plan
        if not (self.use_stack and self.training and self.JK in ("concat", "last") and P is not None and x.is_cuda):
            return None
        pe_zero = data.pe_attr is None or data.pe_attr.numel() == 0 or _all_zero(data.pe_attr)
        norms = [n.module for n in self.norms]
        if data.edge_attr.dim() != 2 or data.edge_attr.size(1) != self.K:
            return None
        plan, _ = get_plan(data.edge_index, data.edge_attr, x.size(0))
        plan.n_dev = getattr(data, "n_dev", None)
        if not stack_applicable(self.gnns, norms, x, P, plan, pe_zero, self.dropout.p):
            if plan.n_dev is not None:
                raise RuntimeError("padded-capacity batches (Batch.n_dev) need the fused layer stack")
            return None
        Hn = kpginplus_stack(self.gnns, norms, x, P, plan, self.residual)      # [N, L+1, H], slot L-j = h_j
        lin = self.output_proj[0]
        if self.JK == "last":
            return self._tail(Hn[:, 0], pre_activation)
        H, L1 = self.hidden_size, self.num_layer + 1
        # JK concat (GNNs.py:455): Hn viewed [N, (L+1)H] holds the layer outputs newest first, so the projection's
        # column blocks are flipped instead of the activations
        Wf = lin.weight.view(lin.out_features, L1, H).flip(1).reshape(lin.out_features, L1 * H)
        rep = _SplitKLinearFn.apply(Hn.view(x.size(0), L1 * H), Wf, lin.bias)
        if pre_activation:
            return rep
        for m in list(self.output_proj)[1:]:
            rep = m(rep)
        return rep


class KPGNNPlusRegressor(nn.Module):
    """GraphRegression (models/GraphRegression.py:10-51) with sum pooling over KPGNNPlusBackbone."""

    def __init__(self, **kw):
        super().__init__()
        self.embedding_model = KPGNNPlusBackbone(**kw)
        self.regressor = nn.Linear(self.embedding_model.hidden_size, 1)

    def forward(self, data):
        h = self.embedding_model(data)
        return self.regressor(segment_sum(h, data.batch, data.num_graphs)).squeeze()

    def fused_loss(self, data, kind="l1"):
        """loss(self(data), data.y) for the L1 (train_ZINC.py:42) or MSE loss with the ReLU of the output projection, the
        pooling, the regressor and the loss in ONE kernel each way (kp_head_*).  None when the backbone's tail is not
        Linear -> ReLU -> identity dropout (the caller then evaluates the loss the usual way)."""
        from .head import fused_regression_loss
        rep = self.embedding_model(data, pre_activation=True)
        if rep is None:
            return None
        return fused_regression_loss(rep, self.regressor.weight, self.regressor.bias, data.y, data.batch,
                                     data.num_graphs, mean=False, kind=kind, n_dev=getattr(data, "n_dev", None))


def zinc_kpginplus(K=8, num_layer=8, hidden=104, combine="geometric"):
    """train_ZINC.py defaults: hidden 104, K 8, 8 layers, max_pe_num 50, 3 edge types, max_edge_count 50,
    max_hop_num 6, max_distance_count 50, JK concat, BatchNorm, sum pooling, residual (README.md:127)."""
    return KPGNNPlusRegressor(num_layer=num_layer, hidden_size=hidden, K=K, input_size=21, num_hop1_edge=3,
                              max_pe_num=50, max_edge_count=50, max_hop_num=6, max_distance_count=50,
                              combine=combine, JK="concat", residual=True, drop_prob=0.0)


def l1_loss(score, y):
    """train_ZINC.py:42"""
    return (score.squeeze() - y.squeeze()).abs().mean()
