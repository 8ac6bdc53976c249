"""Single-launch Adam (csrc/adam.cu, kp_adam_step) with torch.optim.Adam's update rule and defaults
(train_ZINC.py:244: `torch.optim.Adam(model.parameters(), lr=...)`): no amsgrad, no weight decay.

Caller-side glue of the training step, CUDA-graph capturable: moments live in two flat buffers,
the step counter on the device, and the per-tensor pointer table is a pinned host array copied to the device
whenever a gradient tensor's address changes (inside a captured graph: one 4 KB copy node, replayed every step).
"""
import ctypes as C
import struct

import torch

from . import _lib

CHUNK = 1024


class FusedAdam(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params):
            raise _lib.KpError("FusedAdam needs contiguous fp32 CUDA parameters")
        self.lr, self.betas, self.eps = lr, betas, eps
        dev = self.params[0].device
        total = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.state = torch.zeros(2, dtype=torch.int32, device=dev)
        self.offsets, o = [], 0
        for p in self.params:
            self.offsets.append(o)
            o += (p.numel() + 3) // 4 * 4
        chunks = []
        for i, p in enumerate(self.params):
            chunks += [(i, c) for c in range(0, p.numel(), CHUNK)]
        self.nchunks = len(chunks)
        self.chunks = torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous()
        self.table_host = torch.empty(len(self.params) * 40, dtype=torch.uint8, pin_memory=True)
        self.table_dev = torch.empty(len(self.params) * 40, dtype=torch.uint8, device=dev)
        self.key = None

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def exp_avg(self, i):
        p = self.params[i]
        return self.m[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p)

    def exp_avg_sq(self, i):
        p = self.params[i]
        return self.v[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p)

    def _grads(self):
        gs = []
        for p in self.params:
            g = p.grad
            if g is None:
                gs.append(None)
                continue
            if not g.is_contiguous():
                g = g.contiguous()
                p.grad = g
            gs.append(g)
        return gs

    def step(self):
        lib = _lib.lib()
        gs = self._grads()
        key = tuple(g.data_ptr() if g is not None else 0 for g in gs) + tuple(p.data_ptr() for p in self.params)
        dev = self.params[0].device
        if key != self.key:
            if not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(dev).synchronize()     # a previous table copy may still read the pinned buffer
            buf = bytearray()
            mb, vb = self.m.data_ptr(), self.v.data_ptr()
            for p, g, off in zip(self.params, gs, self.offsets):
                n = p.numel() if g is not None else 0            # parameters without a gradient are skipped
                buf += struct.pack("<QQQQii", p.data_ptr(), g.data_ptr() if g is not None else 0, mb + 4 * off,
                                   vb + 4 * off, n, 0)
            self.table_host.copy_(torch.frombuffer(buf, dtype=torch.uint8))
            self.table_dev.copy_(self.table_host, non_blocking=True)
            self.key = key
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.kp_adam_step(self.table_dev.data_ptr(), self.chunks.data_ptr(), self.nchunks, self.lr,
                                    self.betas[0], self.betas[1], self.eps, self.state.data_ptr(), st), "kp_adam_step")
