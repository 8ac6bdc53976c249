"""Fused peripheral-attribute encoder: SURVEY.md section 8(f)-1, the stage right before the K-hop layers.

The reference turns the integer peripheral attributes into `peripheral_attr [N,K,H]` with two
`FeatureConcatEncoder`s (embedding lookups -> concat -> Linear) and a sum over the edge-type slots
(models/GNNs.py:393-400, layers/feature_encoder.py:37-67), materialising [N,K,c,2H] / [N,K,c,H] tensors and a
sort-based embedding backward per table.  Because lookup -> concat -> Linear is linear in the table rows,
        cat_i(E_i[x_i]) W^T + b  =  sum_i (E_i W_i^T)[x_i] + b ,
the whole stage is  P[n,k,:] = sum_s M[slot_off[s] + idx[n,k,s], :]  with M the (tiny) folded, gate-scaled tables.
The folding stays in PyTorch (a handful of [51,H]x[H,H] GEMMs); the gather-sum and its deterministic gradient
are the kp_table_sum_* kernels (include/kpgnn.h).  Parameters stay in the reference's modules, so state_dicts
are unchanged.
"""
import ctypes as C

import torch

from . import _lib


def _ranges(table_sizes, slots_per_table, d, max_bytes=24 * 1024, max_ranges=16):
    """Partition consecutive tables into ranges whose rows fit a shared-memory sub-table and whose slots fit the
    gather-sum kernel's lane group (G = 4 lanes for d <= 16, then the next power of two >= d / 4)."""
    G = 4
    while G < d // 4:
        G <<= 1
    slot_b, row_b = [0], [0]
    rows = slots = 0
    cur = cur_slots = 0
    for n, s in zip(table_sizes, slots_per_table):
        if s > G:
            raise _lib.KpError("a table is read through %d slots but rows of %d floats give lane groups of %d" % (s, d, G))
        if cur and ((cur + n) * d * 4 > max_bytes or cur_slots + s > G):
            slot_b.append(slots)
            row_b.append(rows)
            cur = cur_slots = 0
        cur += n
        cur_slots += s
        rows += n
        slots += s
    slot_b.append(slots)
    row_b.append(rows)
    if len(slot_b) - 1 > max_ranges:
        raise _lib.KpError("too many table ranges (%d)" % (len(slot_b) - 1))
    return slot_b, row_b


class _TableSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, idx, slot_off, range_slot, range_row):
        lib = _lib.lib()
        if not table.is_cuda:
            raise _lib.KpError("kpgnn_b200 runs on CUDA tensors only (no CPU fallback)")
        table = table.contiguous()
        R, S = idx.shape
        d = table.size(1)
        desc = _lib.TsumDesc()
        desc.R, desc.S, desc.d, desc.table_rows = R, S, d, table.size(0)
        desc.idx = idx.data_ptr()
        for i, o in enumerate(slot_off):
            desc.slot_off[i] = o
        desc.num_ranges = len(range_slot) - 1
        for i, (a, b) in enumerate(zip(range_slot, range_row)):
            desc.range_slot[i], desc.range_row[i] = a, b
        out = torch.empty((R, d), dtype=torch.float32, device=table.device)
        st = C.c_void_p(torch.cuda.current_stream(table.device).cuda_stream)
        _lib.check(lib.kp_table_sum_forward(C.byref(desc), table.data_ptr(), out.data_ptr(), st),
                   "kp_table_sum_forward")
        ctx.desc, ctx.idx, ctx.shape = desc, idx, table.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        dout = dout.contiguous()
        if dout.data_ptr() % 16:
            dout = dout.clone()
        dtab = torch.empty(ctx.shape, dtype=torch.float32, device=dout.device)
        nb = C.c_size_t(0)
        _lib.check(lib.kp_table_sum_backward_workspace_bytes(C.byref(ctx.desc), C.byref(nb)), "tsum ws")
        ws = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=dout.device)
        st = C.c_void_p(torch.cuda.current_stream(dout.device).cuda_stream)
        _lib.check(lib.kp_table_sum_backward(C.byref(ctx.desc), dout.data_ptr(), dtab.data_ptr(), ws.data_ptr(),
                                             ws.numel(), st), "kp_table_sum_backward")
        return dtab, None, None, None, None


def peripheral_index(peripheral_edge_attr, peripheral_configuration_attr):
    """[N,K,c,2] and [N,K,h+1] int64 -> one [N*K, 2c + h+1 + 1] int64 matrix, slots ordered by table
    (edge feature 0 x c, edge feature 1 x c, configuration columns, constant bias slot = 0)."""
    N, K = peripheral_edge_attr.shape[:2]
    e = peripheral_edge_attr.reshape(N * K, -1, 2)
    cols = [e[:, :, 0], e[:, :, 1], peripheral_configuration_attr.reshape(N * K, -1),
            torch.zeros((N * K, 1), dtype=torch.int64, device=e.device)]
    return torch.cat(cols, dim=1).contiguous()


import os as _os
_FOLD_KERNEL = _os.environ.get("KP_FOLD", "1") != "0"


class _Fold(torch.autograd.Function):
    """kp_fold_forward / kp_fold_backward (include/kpgnn.h): the folded, gate-scaled lookup table of both encoders.
    Inputs: pew, pcw (raw gates), We, be, Wc, bc (the two Linear layers), then the embedding weights in table order."""

    @staticmethod
    def _desc(pew, pcw, We, be, Wc, bc, embs, n_edge, c_slots, gate_act):
        H_out, H_in = We.size(0), embs[0].size(1)
        d = _lib.FoldDesc()
        d.T, d.H_in, d.H_out, d.gate_act = len(embs), H_in, H_out, gate_act
        off = 0
        for i, e in enumerate(embs):
            edge = i < n_edge
            W = We if edge else Wc
            j = i if edge else i - n_edge
            d.E[i], d.rows[i], d.gate[i] = e.data_ptr(), e.size(0), 0 if edge else 1
            d.W[i], d.w_stride[i] = W.data_ptr() + 4 * j * H_in, W.size(1)
            d.row_off[i] = off
            off += e.size(0)
        d.row_off[len(embs)] = off
        d.gate_raw[0], d.gate_raw[1] = pew.data_ptr(), pcw.data_ptr()
        d.bias[0], d.bias[1] = be.data_ptr(), bc.data_ptr()
        d.bias_mult[0], d.bias_mult[1] = float(c_slots), 1.0
        return d, off + 1

    @staticmethod
    def forward(ctx, n_edge, c_slots, gate_act, pew, pcw, We, be, Wc, bc, *embs):
        lib = _lib.lib()
        ts = [t.detach().contiguous() for t in (pew, pcw, We, be, Wc, bc) + tuple(embs)]
        desc, total = _Fold._desc(*ts[:6], ts[6:], n_edge, c_slots, gate_act)
        table = torch.empty((total, We.size(0)), dtype=torch.float32, device=We.device)
        st = C.c_void_p(torch.cuda.current_stream(We.device).cuda_stream)
        _lib.check(lib.kp_fold_forward(C.byref(desc), table.data_ptr(), st), "kp_fold_forward")
        ctx.ts, ctx.cfg = ts, (n_edge, c_slots, gate_act)
        return table

    @staticmethod
    def backward(ctx, dtable):
        lib = _lib.lib()
        ts = ctx.ts
        n_edge, c_slots, gate_act = ctx.cfg
        pew, pcw, We, be, Wc, bc = ts[:6]
        embs = ts[6:]
        dev = We.device
        desc, _ = _Fold._desc(pew, pcw, We, be, Wc, bc, embs, n_edge, c_slots, gate_act)
        dtable = dtable.contiguous()
        g = _lib.FoldGrads()
        dWe, dWc = torch.empty_like(We), torch.empty_like(Wc)
        dbe, dbc = torch.empty_like(be), torch.empty_like(bc)
        dpe, dpc = torch.empty_like(pew), torch.empty_like(pcw)
        dembs = [torch.empty_like(e) for e in embs]
        H_in = embs[0].size(1)
        for i, de in enumerate(dembs):
            edge = i < n_edge
            g.dE[i] = de.data_ptr()
            g.dW[i] = (dWe if edge else dWc).data_ptr() + 4 * (i if edge else i - n_edge) * H_in
        g.dbias[0], g.dbias[1] = dbe.data_ptr(), dbc.data_ptr()
        g.dgate_raw[0], g.dgate_raw[1] = dpe.data_ptr(), dpc.data_ptr()
        ws = _lib.fold_workspace(dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.kp_fold_backward(C.byref(desc), dtable.data_ptr(), C.byref(g), ws.data_ptr(), ws.numel(), st),
                   "kp_fold_backward")
        return (None, None, None, dpe, dpc, dWe, dbe, dWc, dbc) + tuple(dembs)


def fused_peripheral_attr(edge_enc, cfg_enc, pew, pcw, idx, N, K, c_slots, gate="tanh"):
    """P [N,K,H] = g(pew) * edge_enc(edge_attr).sum(-2) + g(pcw) * cfg_enc(cfg_attr), computed as one gather-sum over
    the folded tables.  edge_enc / cfg_enc are FeatureConcatEncoder modules (2 and h+1 tables); pew / pcw the RAW gate
    parameters, g = tanh in GNNPlus (GNNs.py:396), sigmoid in GNN / GNNPrime (GNNs.py:175); idx from `peripheral_index`.
    The fold (nine small products forward, twenty-seven backward) is one kernel each way (kp_fold_*)."""
    H = edge_enc.proj.out_features
    embs = [e.weight for e in edge_enc.embedding_list] + [e.weight for e in cfg_enc.embedding_list]
    nc = len(cfg_enc.embedding_list)
    same = all(e.size(1) == embs[0].size(1) for e in embs) and edge_enc.proj.in_features == 2 * embs[0].size(1)
    if not same or len(embs) > 16 or H > 256 or embs[0].size(1) > 256:
        raise _lib.KpError("fused_peripheral_attr: unsupported encoder shapes")
    if _FOLD_KERNEL:
        table = _Fold.apply(len(edge_enc.embedding_list), c_slots, 0 if gate == "tanh" else 1, pew, pcw,
                            edge_enc.proj.weight, edge_enc.proj.bias, cfg_enc.proj.weight, cfg_enc.proj.bias, *embs)
    else:                      # the same fold as ~50 library launches (A/B measurements only)
        Hi = embs[0].size(1)
        ge, gc = (torch.tanh(pew), torch.tanh(pcw)) if gate == "tanh" else (torch.sigmoid(pew), torch.sigmoid(pcw))
        We, Wc = edge_enc.proj.weight, cfg_enc.proj.weight
        tabs = [ge * (edge_enc.embedding_list[i].weight @ We[:, i * Hi:(i + 1) * Hi].t()) for i in range(2)]
        Ec = torch.stack([e.weight for e in cfg_enc.embedding_list])
        tabs.append((gc * torch.bmm(Ec, Wc.view(H, nc, Hi).permute(1, 2, 0))).reshape(-1, H))
        tabs.append((c_slots * ge * edge_enc.proj.bias + gc * cfg_enc.proj.bias).view(1, H))
        table = torch.cat(tabs, dim=0)
    sizes = [edge_enc.embedding_list[0].num_embeddings, edge_enc.embedding_list[1].num_embeddings] + \
            [e.num_embeddings for e in cfg_enc.embedding_list] + [1]
    slots = [c_slots, c_slots] + [1] * nc + [1]
    offs, o = [], 0
    for n, s_ in zip(sizes, slots):
        offs += [o] * s_
        o += n
    range_slot, range_row = _ranges(sizes, slots, H)
    out = _TableSum.apply(table, idx, offs, range_slot, range_row)
    return out.view(N, K, H)
