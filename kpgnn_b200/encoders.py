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
    """Partition consecutive tables into ranges whose rows fit a shared-memory sub-table."""
    slot_b, row_b = [0], [0]
    rows = slots = 0
    cur = 0
    for n, s in zip(table_sizes, slots_per_table):
        if cur and (cur + n) * d * 4 > max_bytes:
            slot_b.append(slots)
            row_b.append(rows)
            cur = 0
        cur += n
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


def fused_peripheral_attr(edge_enc, cfg_enc, gate_e, gate_c, idx, N, K, c_slots):
    """P [N,K,H] = gate_e * edge_enc(edge_attr).sum(-2) + gate_c * cfg_enc(cfg_attr), computed as one gather-sum.
    edge_enc / cfg_enc are FeatureConcatEncoder modules (2 and h+1 tables); gate_* are the already-squashed
    scalars (tanh(pew) in GNNPlus, sigmoid in GNN/GNNPrime); idx from `peripheral_index`."""
    H = edge_enc.proj.out_features
    We = edge_enc.proj.weight                    # [H, 2H]
    Wc = cfg_enc.proj.weight                     # [H, (h+1)H]
    nc = len(cfg_enc.embedding_list)
    tabs = [gate_e * (edge_enc.embedding_list[i].weight @ We[:, i * H:(i + 1) * H].t()) for i in range(2)]
    Ec = torch.stack([e.weight for e in cfg_enc.embedding_list])            # [h+1, R, H]
    Wc3 = Wc.view(H, nc, H).permute(1, 2, 0)                                # [h+1, H_in, H_out]
    tabs.append((gate_c * torch.bmm(Ec, Wc3)).reshape(-1, H))
    tabs.append((c_slots * gate_e * edge_enc.proj.bias + gate_c * cfg_enc.proj.bias).view(1, H))
    table = torch.cat(tabs, dim=0)
    sizes = [edge_enc.embedding_list[0].num_embeddings, edge_enc.embedding_list[1].num_embeddings] + \
            [e.num_embeddings for e in cfg_enc.embedding_list] + [1]
    slots = [c_slots, c_slots] + [1] * nc + [1]
    offs, o = [], 0
    for n, s in zip(sizes, slots):
        offs += [o] * s
        o += n
    range_slot, range_row = _ranges(sizes, slots, H)
    out = _TableSum.apply(table, idx, offs, range_slot, range_row)
    return out.view(N, K, H)
