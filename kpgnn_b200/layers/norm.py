"""Training-mode BatchNorm1d (+ fused ReLU) on one sm_100a kernel each way (csrc/bn.cu, kp_bn_forward/backward).

The KP-GIN+ layer's MLP is Linear-BN-ReLU-Linear-BN-ReLU (layers/KPGINplus.py:25-30) and the backbone normalises
every layer's output with a BatchNorm (models/GNNs.py:324-325,430).  At molecule-batch sizes (a few thousand nodes,
~100 channels) PyTorch runs each BatchNorm as three kernels forward and two backward plus two for the ReLU; after
the aggregation itself was fused that glue was ~24 % of the training step (profiles/r1j_step_launches.csv).

`FusedBatchNorm1d` is an `nn.BatchNorm1d` (same parameters, buffers and state_dict keys).  The kernels cover the
training-mode forward/backward for N <= kp_bn_max_rows() rows; eval mode, CPU tensors, larger batches and exotic
settings (momentum=None, no affine) take torch's own batch_norm -- this module is caller-side glue,
not the K-hop path, so that fallback is the framework's operator, not an oracle.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


class _BNTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, nbt, eps, momentum, relu):
        lib = _lib.lib()
        x = x.contiguous()
        N, Cn = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(Cn, dtype=torch.float32, device=x.device)
        invstd = torch.empty(Cn, dtype=torch.float32, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.kp_bn_forward(x.data_ptr(), N, Cn, weight.data_ptr(), bias.data_ptr(), eps, momentum,
                                     running_mean.data_ptr() if running_mean is not None else None,
                                     running_var.data_ptr() if running_var is not None else None,
                                     nbt.data_ptr() if nbt is not None else None, 1 if relu else 0,
                                     y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), st), "kp_bn_forward")
        ctx.save_for_backward(x, weight, bias, mean, invstd)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.lib()
        x, weight, bias, mean, invstd = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.data_ptr() % 16:
            dy = dy.clone()
        N, Cn = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(bias)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.kp_bn_backward(x.data_ptr(), dy.data_ptr(), N, Cn, weight.data_ptr(), bias.data_ptr(),
                                      mean.data_ptr(), invstd.data_ptr(), 1 if ctx.relu else 0, dx.data_ptr(),
                                      dgamma.data_ptr(), dbeta.data_ptr(), st), "kp_bn_backward")
        return dx, dgamma, dbeta, None, None, None, None, None, None


class FusedBatchNorm1d(nn.BatchNorm1d):
    """nn.BatchNorm1d whose training step on CUDA is one kernel; `relu=True` folds the following ReLU in."""

    def __init__(self, num_features, relu=False, **kw):
        super().__init__(num_features, **kw)
        self.fuse_relu = relu

    def _kernel_ok(self, x):
        return (self.training and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and self.affine
                and self.momentum is not None and x.size(1) % 4 == 0 and 1 < x.size(0) <= _lib.lib().kp_bn_max_rows()
                and x.data_ptr() % 16 == 0)

    def forward(self, x):
        if self._kernel_ok(x):
            return _BNTrain.apply(x, self.weight, self.bias,
                                  self.running_mean if self.track_running_stats else None,
                                  self.running_var if self.track_running_stats else None,
                                  self.num_batches_tracked if self.track_running_stats else None,
                                  self.eps, self.momentum, self.fuse_relu)
        y = super().forward(x)
        return F.relu(y) if self.fuse_relu else y

    def extra_repr(self):
        return super().extra_repr() + ", fused_relu=%s" % self.fuse_relu
