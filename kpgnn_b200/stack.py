"""The L KP-GIN+ layers of the backbone as ONE autograd node over a layer-history buffer.

The reference builds every layer's input with `torch.stack` of the last k layer outputs (models/GNNs.py:413-418),
normalises, adds the residual and finally concatenates all layer outputs (GNNs.py:430-438,455-456).  Run op by op
that is, per layer, a cat kernel forward and -- because each layer output feeds up to K later stacks, the residual and
the JK concat -- up to ten gradient-accumulation kernels backward; the shared peripheral tensor P[:, :k] adds a
zero-padded slice and an accumulation per layer (profiles/r1z_step_kineto.txt: ~110 add/copy/cat/fill launches of the
245 per step).

Here the layer outputs live in one tensor `Hn [N, L+1, H]`, slot L-j = h_j (newest first):
  * layer l's input `[N, k, H]` is the strided view Hn[:, L-l : L-l+k, :] -- the aggregation kernels take node / hop
    strides, so there is no stack;
  * the dense-block kernel writes norm(mlp(.)) + residual straight into slot L-l-1 and reads the residual from slot L-l;
  * `Hn.view(N, (L+1)H)` IS the JK concat (with the column blocks in reverse layer order: the caller flips the output
    projection's weight blocks instead of the activations);
  * backward walks the layers in reverse over ONE gradient buffer of the same shape: the dense-block backward reads its
    dOut from slot L-l-1 and accumulates the residual's gradient into slot L-l in-kernel, the aggregation's dX is added
    to slots L-l .. L-l+k-1, and the gradient of P is ONE kernel over the L aggregation-output gradients
    (kp_peripheral_grad) instead of L padded slices.
Same kernels, same arithmetic and summation order inside every kernel as the per-layer path; only the order in which
autograd would have added the per-consumer gradients of a layer output differs (fp32 rounding, ~1e-7 relative).
"""
import ctypes as C

import torch

from . import _lib
from .ops import _make_desc, want_blocks, ACT_GELU
from .layers.combine import GeometricCombine
from .layers.dense_block import _bn_ok


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _KPGINPlusStack(torch.autograd.Function):
    """forward(x0 [N,H], P [N,K,H], *params, cfg) -> Hn [N, L+1, H]; params are, per layer, in `_layer_params` order."""

    @staticmethod
    def forward(ctx, x0, P, cfg, *params):
        lib = _lib.lib()
        plan, layers, norms, residual = cfg
        L = len(layers)
        N, H = x0.shape
        dev = x0.device
        K = P.size(1)
        st = _stream(dev)
        Pc = P.detach()
        if Pc.stride(-1) != 1:
            Pc = Pc.contiguous()
        Hn = torch.empty((N, L + 1, H), dtype=torch.float32, device=dev)
        Hn[:, L].copy_(x0.detach())
        want_blocks(plan, K, H, True)
        hs = (L + 1) * H
        saved = []
        # GeometricCombine weights of every layer in one launch (combine.py:51-58)
        thetas = [None] * L
        tb = _lib.ThetaBatch()
        tb.d = H
        pi, nb_ = 0, 0
        for l, layer in enumerate(layers):
            if layer.K > 1:
                thetas[l] = torch.empty((layer.K, H), dtype=torch.float32, device=dev)
                tb.alphas[nb_], tb.theta[nb_], tb.k[nb_] = (params[pi + 12].data_ptr(), thetas[l].data_ptr(), layer.K)
                nb_ += 1
            pi += _num_params(layer)
        tb.L = nb_
        _lib.check(lib.kp_geometric_theta_forward_batched(C.byref(tb), st), "kp_geometric_theta_forward_batched")
        pi = 0
        for l, layer in enumerate(layers):
            k = layer.K
            np_l = _num_params(layer)
            W1, b1, g1, be1, W2, b2, g2, be2, g3, be3, T0 = (t.detach() for t in params[pi:pi + 11])
            Tk = params[pi + 11].detach() if k > 1 else None
            alphas = params[pi + 12].detach() if k > 1 else None
            pi += np_l
            plan.check_tables(T0.size(0), Tk.size(0) if Tk is not None else 0, k)
            theta = thetas[l]
            xs = Hn[:, L - l:L - l + k, :]
            adesc = _make_desc(plan, k, xs, Pc[:, :k], T0.contiguous(), Tk.contiguous() if Tk is not None else None,
                               theta, None, ACT_GELU, k > 1, False, False)
            agg = torch.empty((N, H), dtype=torch.float32, device=dev)            # fuse: [N,H]; k == 1: [N,1,H]
            _lib.check(lib.kp_agg_forward(C.byref(adesc), agg.data_ptr(), st), "kp_agg_forward")
            bn1, bn2, bn3 = layer.mlp[1], layer.mlp[4], norms[l]
            d = _lib.DenseDesc()
            d.N, d.Cin, d.Cout = N, H, H
            W1c, W2c = W1.contiguous(), W2.contiguous()
            d.X, d.W1, d.b1, d.g1, d.be1 = agg.data_ptr(), W1c.data_ptr(), b1.data_ptr(), g1.data_ptr(), be1.data_ptr()
            d.W2, d.b2, d.g2, d.be2 = W2c.data_ptr(), b2.data_ptr(), g2.data_ptr(), be2.data_ptr()
            d.g3, d.be3 = g3.data_ptr(), be3.data_ptr()
            d.eps1, d.eps2, d.eps3 = bn1.eps, bn2.eps, bn3.eps
            d.mom1, d.mom2, d.mom3 = bn1.momentum, bn2.momentum, bn3.momentum
            d.rm1, d.rv1, d.nbt1 = (bn1.running_mean.data_ptr(), bn1.running_var.data_ptr(),
                                    bn1.num_batches_tracked.data_ptr())
            d.rm2, d.rv2, d.nbt2 = (bn2.running_mean.data_ptr(), bn2.running_var.data_ptr(),
                                    bn2.num_batches_tracked.data_ptr())
            d.rm3, d.rv3, d.nbt3 = (bn3.running_mean.data_ptr(), bn3.running_var.data_ptr(),
                                    bn3.num_batches_tracked.data_ptr())
            keep = torch.empty((3, N, H), dtype=torch.float32, device=dev)
            stats = torch.empty((6, H), dtype=torch.float32, device=dev)
            d.Y1, d.Y2, d.Z2, d.stats = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), stats.data_ptr()
            out = Hn[:, L - l - 1]
            d.barrier = _lib.barrier_state(dev).data_ptr()
            if plan.n_dev is not None:          # N is a padded capacity: the batch's row count lives on the device
                d.n_dev = plan.n_dev.data_ptr()
            d.out_stride = hs
            if residual:
                d.R, d.r_stride = Hn[:, L - l].data_ptr(), hs
            fb, bb = C.c_size_t(0), C.c_size_t(0)
            _lib.check(lib.kp_dense_block_workspace_bytes(C.byref(d), C.byref(fb), C.byref(bb)), "kp_dense_block ws")
            ws = torch.empty(fb.value, dtype=torch.uint8, device=dev)
            _lib.check(lib.kp_dense_block_forward(C.byref(d), out.data_ptr(), ws.data_ptr(), ws.numel(), st),
                       "kp_dense_block_forward")
            saved.append((adesc, d, bb.value, k, theta, alphas, agg, keep, stats, (W1c, W2c, T0, Tk)))
        # a non-autograd alias keeps the storage alive for the saved descriptors without the node -> output -> grad_fn
        # reference cycle that storing the output itself would create (freed by refcount, not by the cyclic GC)
        ctx.saved, ctx.Hn, ctx.Pc, ctx.cfg = saved, Hn.detach(), Pc, cfg
        ctx.shape = (N, H, K, L)
        return Hn

    @staticmethod
    def backward(ctx, dHn):
        lib = _lib.lib()
        plan, layers, norms, residual = ctx.cfg
        N, H, K, L = ctx.shape
        dev = dHn.device
        st = _stream(dev)
        G = dHn.contiguous().clone()                   # accumulated in place below
        # leaf gradients (tables, combine weights, dense weights) go to a second stream and overlap the chain of
        # dX-producing kernels; everything they touch is kept alive until the join at the end
        main = torch.cuda.current_stream(dev)
        leaf = _leaf_stream(dev) if _USE_LEAF_STREAM else None
        if leaf is not None:
            leaf.wait_stream(main)
        keep_alive = []
        hs = (L + 1) * H
        grads = [None] * sum(_num_params(layer) for layer in layers)
        offs, o = [], 0
        for layer in layers:
            offs.append(o)
            o += _num_params(layer)
        pg = _lib.PgradDesc()
        pg.N, pg.K, pg.d, pg.L = N, K, H, L
        daggs = []
        for l in range(L - 1, -1, -1):
            adesc, d, bwd_bytes, k, theta, alphas, agg, keep, stats, (W1c, W2c, T0, Tk) = ctx.saved[l]
            dagg = torch.empty((N, H), dtype=torch.float32, device=dev)
            dW1, dW2 = torch.empty_like(W1c), torch.empty_like(W2c)
            dvec = torch.empty((8, H), dtype=torch.float32, device=dev)   # db1 db2 | dg1 dbe1 dg2 dbe2 dg3 dbe3
            d.dout_stride = hs
            if residual:
                d.dR, d.dr_stride = G[:, L - l].data_ptr(), hs
            ws = torch.empty(bwd_bytes, dtype=torch.uint8, device=dev)
            d.leaf_stream = leaf.cuda_stream if leaf is not None else None
            adesc.leaf_stream = leaf.cuda_stream if leaf is not None else None
            _lib.check(lib.kp_dense_block_backward(C.byref(d), G[:, L - l - 1].data_ptr(), dagg.data_ptr(),
                                                   dW1.data_ptr(), dvec[0].data_ptr(), dW2.data_ptr(),
                                                   dvec[1].data_ptr(), dvec[2].data_ptr(), ws.data_ptr(), ws.numel(),
                                                   st), "kp_dense_block_backward")
            dT0 = torch.empty_like(T0)
            dTk = torch.empty_like(Tk) if Tk is not None else None
            dal = torch.empty_like(alphas) if theta is not None else None
            if theta is not None:       # dtheta reduction fused with GeometricCombine's backward
                adesc.geo_alphas, adesc.geo_dalphas = alphas.data_ptr(), dal.data_ptr()
            nb = C.c_size_t(0)
            _lib.check(lib.kp_agg_backward_workspace_bytes(C.byref(adesc), C.byref(nb)), "kp_agg_backward ws")
            ws2 = torch.empty(max(nb.value, 1), dtype=torch.uint8, device=dev)

            def agg_bwd(dx_ptr):
                return lib.kp_agg_backward(C.byref(adesc), dagg.data_ptr(), dx_ptr, None, dT0.data_ptr(),
                                           dTk.data_ptr() if dTk is not None else None,
                                           None, None, ws2.data_ptr(), ws2.numel(), st)
            # dX is accumulated straight into the history gradient (slots L-l .. L-l+k-1) by the gather kernel ...
            adesc.dx_node_stride, adesc.dx_hop_stride, adesc.dx_accumulate = hs, H, 1
            rc = agg_bwd(G[:, L - l].data_ptr())
            if rc == 3:                 # ... unless another kernel family has to serve this call: temporary + add
                adesc.dx_node_stride, adesc.dx_hop_stride, adesc.dx_accumulate = 0, 0, 0
                dX = torch.empty((N, k, H), dtype=torch.float32, device=dev)
                _lib.check(agg_bwd(dX.data_ptr()), "kp_agg_backward")
                G[:, L - l:L - l + k].add_(dX)
            else:
                _lib.check(rc, "kp_agg_backward")
            keep_alive += [ws, ws2, dW1, dW2, dvec, dT0, dTk, dal]
            pg.dagg[l], pg.k[l] = dagg.data_ptr(), k
            pg.theta[l] = theta.data_ptr() if theta is not None else None
            daggs.append(dagg)
            g = [dW1, dvec[0], dvec[2], dvec[3], dW2, dvec[1], dvec[4], dvec[5], dvec[6], dvec[7], dT0]
            if k > 1:
                g += [dTk, dal]
            grads[offs[l]:offs[l] + len(g)] = g
        dP = None
        if ctx.needs_input_grad[1]:
            dP = torch.empty((N, K, H), dtype=torch.float32, device=dev)
            _lib.check(lib.kp_peripheral_grad(C.byref(pg), dP.data_ptr(), st), "kp_peripheral_grad")
        if leaf is not None:
            main.wait_stream(leaf)                     # join: leaf gradients are complete for whoever runs next
        del keep_alive
        dx0 = G[:, L] if ctx.needs_input_grad[0] else None
        return (dx0, dP, None) + tuple(grads)


_USE_LEAF_STREAM = True
_LEAF = {}


def _leaf_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _LEAF.get(key)
    if s is None:
        s = torch.cuda.Stream(dev)
        _LEAF[key] = s
    return s


def _num_params(layer):
    return 13 if layer.K > 1 else 11


def _layer_params(layer, norm):
    """W1 b1 g1 be1 W2 b2 g2 be2 g3 be3 T0 [Tk alphas]"""
    m = layer.mlp
    p = [m[0].weight, m[0].bias, m[1].weight, m[1].bias, m[3].weight, m[3].bias, m[4].weight, m[4].bias,
         norm.weight, norm.bias, layer.hop1_edge_emb.weight]
    if layer.K > 1:
        p += [layer.hopk_edge_emb.weight, layer.combine.alphas]
    return p


def stack_applicable(layers, norms, x0, P, plan, pe_attr_zero, dropout_p):
    """True when the whole stack can run as one node: training mode, geometric combine, BatchNorm everywhere, no
    dropout between layers, pe_attr == 0 (what the reference's extractor emits), everything fp32 on one CUDA device."""
    from .layers.KPGINplus import KPGINPlusConv
    from .layers import dense_block as _db
    if _db._DISABLED:
        return False
    if not (torch.is_tensor(x0) and x0.is_cuda and x0.dim() == 2 and x0.dtype == torch.float32 and P is not None
            and P.dtype == torch.float32 and P.dim() == 3 and pe_attr_zero and dropout_p == 0.0):
        return False
    N, H = x0.shape
    if H % 4 or H > 128 or N < 2 or N > _lib.lib().kp_dense_block_max_rows(H, H) or P.size(0) != N or P.size(2) != H:
        return False
    for l, (layer, norm) in enumerate(zip(layers, norms)):
        if not isinstance(layer, KPGINPlusConv) or not layer.training or layer.K != min(l + 1, P.size(1)):
            return False
        if layer.K > 1 and not isinstance(layer.combine, GeometricCombine):
            return False
        m = layer.mlp
        if not (_bn_ok(m[1]) and _bn_ok(m[4]) and _bn_ok(norm) and m[0].bias is not None and m[3].bias is not None):
            return False
        if m[0].in_features != H or m[0].out_features != H or m[3].out_features != H or norm.num_features != H:
            return False
    return True


def kpginplus_stack(layers, norms, x0, P, plan, residual=True):
    """Runs the L layers; returns Hn [N, L+1, H] with Hn[:, L-j] = h_j (h_0 = x0)."""
    params = []
    for layer, norm in zip(layers, norms):
        params += _layer_params(layer, norm)
    return _KPGINPlusStack.apply(x0, P, (plan, list(layers), list(norms), bool(residual)), *params)
