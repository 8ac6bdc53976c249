"""Fused dense block of a KP-GIN+ layer (csrc/dense.cu, kp_dense_block_forward/backward): in training mode

    mlp(h)            = ReLU(BN2(Linear2(ReLU(BN1(Linear1(h))))))            layers/KPGINplus.py:25-30,78
    norm(mlp(h)) + r  = the backbone's BatchNorm and residual add            models/GNNs.py:430-438

run as ONE persistent kernel forward and ONE backward instead of ~45 library / elementwise launches per layer.
The modules keep their parameters, buffers and state_dict keys (`mlp.0.weight`, `mlp.1.running_mean`, ...): this
function only reads them.  It is caller-side glue around the K-hop path; whenever its
preconditions do not hold (eval mode, CPU tensors, more rows than fit one slab per SM, no running statistics,
momentum=None) `fused_dense_block` returns None and the caller runs the modules one by one.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib


import os as _os
_DISABLED = _os.environ.get("KP_NO_DENSE") == "1"        # A/B switch: run the modules one by one


class _DenseBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, g1, be1, W2, b2, g2, be2, g3, be3, res, bns, n_dev=None):
        lib = _lib.lib()
        bn1, bn2, bn3 = bns
        x = x.contiguous()
        N, Cin = x.shape
        Cout = W1.size(0)
        dev = x.device
        d = _lib.DenseDesc()
        d.N, d.Cin, d.Cout = N, Cin, Cout
        W1c, W2c = W1.contiguous(), W2.contiguous()
        d.X, d.W1, d.b1, d.g1, d.be1 = x.data_ptr(), W1c.data_ptr(), b1.data_ptr(), g1.data_ptr(), be1.data_ptr()
        d.W2, d.b2, d.g2, d.be2 = W2c.data_ptr(), b2.data_ptr(), g2.data_ptr(), be2.data_ptr()
        d.g3 = g3.data_ptr() if g3 is not None else None
        d.be3 = be3.data_ptr() if be3 is not None else None
        resc = res.contiguous() if res is not None else None
        d.R = resc.data_ptr() if resc is not None else None
        d.eps1, d.eps2, d.mom1, d.mom2 = bn1.eps, bn2.eps, bn1.momentum, bn2.momentum
        d.rm1, d.rv1, d.nbt1 = bn1.running_mean.data_ptr(), bn1.running_var.data_ptr(), bn1.num_batches_tracked.data_ptr()
        d.rm2, d.rv2, d.nbt2 = bn2.running_mean.data_ptr(), bn2.running_var.data_ptr(), bn2.num_batches_tracked.data_ptr()
        if bn3 is not None:
            d.eps3, d.mom3 = bn3.eps, bn3.momentum
            d.rm3, d.rv3, d.nbt3 = (bn3.running_mean.data_ptr(), bn3.running_var.data_ptr(),
                                    bn3.num_batches_tracked.data_ptr())
        saved = torch.empty((3 if g3 is not None else 2, N, Cout), dtype=torch.float32, device=dev)
        stats = torch.empty((6, Cout), dtype=torch.float32, device=dev)
        d.Y1, d.Y2 = saved[0].data_ptr(), saved[1].data_ptr()
        d.Z2 = saved[2].data_ptr() if g3 is not None else None
        d.stats = stats.data_ptr()
        out = torch.empty((N, Cout), dtype=torch.float32, device=dev)
        d.barrier = _lib.barrier_state(dev).data_ptr()
        if n_dev is not None:
            d.n_dev = n_dev.data_ptr()
        fb, bb = C.c_size_t(0), C.c_size_t(0)
        _lib.check(lib.kp_dense_block_workspace_bytes(C.byref(d), C.byref(fb), C.byref(bb)), "kp_dense_block ws")
        ws = torch.empty(fb.value, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.kp_dense_block_forward(C.byref(d), out.data_ptr(), ws.data_ptr(), ws.numel(), st),
                   "kp_dense_block_forward")
        ctx.desc, ctx.bwd_bytes, ctx.has_bn3, ctx.has_res = d, bb.value, g3 is not None, res is not None
        ctx.keep = (x, W1c, W2c, b1, g1, be1, b2, g2, be2, g3, be3, saved, stats, n_dev)   # owners of the desc's pointers
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        d = ctx.desc
        x, W1c, W2c = ctx.keep[0], ctx.keep[1], ctx.keep[2]
        dev = dout.device
        dout = dout.contiguous()
        if dout.data_ptr() % 16:
            dout = dout.clone()
        Cout = d.Cout
        dX = torch.empty_like(x)
        dW1, dW2 = torch.empty_like(W1c), torch.empty_like(W2c)
        dvec = torch.empty((8, Cout), dtype=torch.float32, device=dev)   # db1, db2, dg1, dbe1, dg2, dbe2, dg3, dbe3
        ws = torch.empty(ctx.bwd_bytes, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.kp_dense_block_backward(C.byref(d), dout.data_ptr(), dX.data_ptr(), dW1.data_ptr(),
                                               dvec[0].data_ptr(), dW2.data_ptr(), dvec[1].data_ptr(),
                                               dvec[2].data_ptr(), ws.data_ptr(), ws.numel(), st),
                   "kp_dense_block_backward")
        g3g = dvec[6] if ctx.has_bn3 else None
        be3g = dvec[7] if ctx.has_bn3 else None
        return (dX, dW1, dvec[0], dvec[2], dvec[3], dW2, dvec[1], dvec[4], dvec[5], g3g, be3g,
                dout if ctx.has_res else None, None, None)


def _bn_ok(bn):
    return (isinstance(bn, nn.BatchNorm1d) and bn.training and bn.affine and bn.track_running_stats
            and bn.momentum is not None and bn.weight is not None and bn.weight.dtype == torch.float32)


def fused_dense_block(x, lin1, bn1, lin2, bn2, bn3=None, residual=None, n_dev=None):
    """Linear-BN-ReLU-Linear-BN-ReLU (+ BatchNorm + residual) in one kernel, or None if not applicable.
    n_dev: optional device int32 scalar -- the number of rows that exist when x.size(0) is a padded capacity
    (kp_dense_desc.n_dev): padding rows are excluded from every statistic and come back as zeros."""
    if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32) or _DISABLED:
        return None
    if not (isinstance(lin1, nn.Linear) and isinstance(lin2, nn.Linear) and lin1.bias is not None
            and lin2.bias is not None and _bn_ok(bn1) and _bn_ok(bn2) and (bn3 is None or _bn_ok(bn3))):
        return None
    Cin, Cout = lin1.in_features, lin1.out_features
    if lin2.in_features != Cout or lin2.out_features != Cout or bn1.num_features != Cout \
            or bn2.num_features != Cout or (bn3 is not None and bn3.num_features != Cout):
        return None
    N = x.size(0)
    if N < 2 or N > _lib.lib().kp_dense_block_max_rows(Cin, Cout) or x.size(1) != Cin:
        return None
    if residual is not None and (residual.shape != (N, Cout) or residual.dtype != torch.float32):
        return None
    if x.data_ptr() % 16 or (residual is not None and residual.is_contiguous() and residual.data_ptr() % 16):
        return None
    return _DenseBlock.apply(x, lin1.weight, lin1.bias, bn1.weight, bn1.bias, lin2.weight, lin2.bias, bn2.weight,
                             bn2.bias, bn3.weight if bn3 is not None else None,
                             bn3.bias if bn3 is not None else None, residual, (bn1, bn2, bn3), n_dev)
