"""Fused graph-regression head (kp_head_forward / kp_head_backward, include/kpgnn.h): ReLU of the backbone's output
projection, per-graph pooling, the Linear(H,1) regressor and the L1 / MSE loss -- models/GNNs.py:276-277,
models/GraphRegression.py:26 and train_ZINC.py:42 -- as one kernel each way."""
import ctypes as C

import torch

from . import _lib

_WS = {}


def _workspace(device, G):
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(device).cuda_stream, int(G))
    t = _WS.get(key)
    if t is None:
        t = torch.zeros(int(_lib.lib().kp_head_workspace_bytes(int(G))), dtype=torch.uint8, device=device)
        _WS[key] = t
    return t


class _FusedRegressionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rep, w, b, y, batch, G, mean, kind, n_dev):
        if not rep.is_cuda:
            raise _lib.KpError("kpgnn_b200 runs on CUDA tensors only (no CPU fallback); got rep on %s" % rep.device)
        rep = rep.detach()
        if rep.dtype != torch.float32 or rep.stride(1) != 1:
            rep = rep.float().contiguous()
        N, H = rep.shape
        wv = w.detach().reshape(-1).contiguous()
        bv = b.detach().reshape(-1).contiguous()
        yv = y.detach().reshape(-1).to(torch.float32).contiguous()
        assert wv.numel() == H and yv.numel() == G and batch.dtype == torch.int64
        d = _lib.HeadDesc()
        d.N, d.H, d.G, d.mean, d.loss_kind = N, H, int(G), 1 if mean else 0, {"l1": 0, "mse": 1}[kind]
        d.rep, d.rep_stride, d.rep_stride_out = rep.data_ptr(), rep.stride(0), H
        d.batch, d.n_dev = batch.data_ptr(), (n_dev.data_ptr() if n_dev is not None else None)
        d.w, d.b, d.y = wv.data_ptr(), bv.data_ptr(), yv.data_ptr()
        dev = rep.device
        pooled = torch.empty((G, H), dtype=torch.float32, device=dev)
        score = torch.empty(G, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = _workspace(dev, G)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().kp_head_forward(C.byref(d), pooled.data_ptr(), score.data_ptr(), loss.data_ptr(),
                                              ws.data_ptr(), ws.numel(), st), "kp_head_forward")
        ctx.desc, ctx.keep = d, (rep, wv, bv, yv, batch, n_dev, pooled, score, ws)
        ctx.shapes = (w.shape, b.shape)
        ctx.mark_non_differentiable(score)
        return loss, score

    @staticmethod
    def backward(ctx, dloss, _dscore):
        rep, wv, bv, yv, batch, n_dev, pooled, score, ws = ctx.keep
        d = ctx.desc
        dev = rep.device
        drep = torch.empty((d.N, d.H), dtype=torch.float32, device=dev)
        dw = torch.empty(d.H, dtype=torch.float32, device=dev)
        db = torch.empty(1, dtype=torch.float32, device=dev)
        dl = dloss.detach().reshape(1).to(torch.float32).contiguous()
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().kp_head_backward(C.byref(d), pooled.data_ptr(), score.data_ptr(), dl.data_ptr(),
                                               drep.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), st),
                   "kp_head_backward")
        return drep, dw.view(ctx.shapes[0]), db.view(ctx.shapes[1]), None, None, None, None, None, None


def fused_regression_loss(rep, w, b, y, batch, num_graphs, mean=False, kind="l1", n_dev=None, return_score=False):
    """rep [N,H]: output of the backbone's projection BEFORE its ReLU; w, b: the Linear(H,1) regressor; returns the scalar
    loss (and the per-graph scores)."""
    loss, score = _FusedRegressionLoss.apply(rep, w, b, y, batch, int(num_graphs), bool(mean), kind, n_dev)
    return (loss, score) if return_score else loss
