"""Compact batch wire format and its device-side unpacking (csrc/wire.cu, kp_wire_unpack).

The reference collates a new batch on the host every step and ships int64 tensors (train_ZINC.py:29-47, PyG
`Batch.from_data_list` + `.to(device)`): 7.5 MB per 128 molecules, a different node / edge count every time.  Here
  * `WireSpec.pack()` writes one batch into a flat pinned byte buffer of FIXED size: int32 node ids, 1- or 2-byte
    attributes, per-graph node offsets, targets (~1.3 MB at the same batch);
  * `DeviceWire.unpack()` widens it on the device into STATIC int64 tensors of a fixed capacity -- the reference's wire
    layout, so every drop-in layer and backbone consumes it unchanged -- padding the tail (masked edges, zero nodes
    with graph id == num_graphs) and writing the batch's node count to `n_dev`.
Static addresses + a device-side row count mean the whole training step can be captured in ONE CUDA graph and replayed
for batches of different sizes (kpgnn_b200/train.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import Batch


def _align(n, a=256):
    return (n + a - 1) // a * a


class WireSpec(object):
    """Capacities and element widths of one stream of batches; fixes the byte layout of the flat buffer."""

    def __init__(self, n_cap, e_cap, num_graphs, K, max_edge_type, max_hop_num, attr_max=255, periph_max=255,
                 x_max=255):
        self.n_cap, self.e_cap, self.G, self.K = int(n_cap), int(e_cap), int(num_graphs), int(K)
        self.met, self.hp1 = int(max_edge_type), int(max_hop_num) + 1
        self.x_bytes = 1 if x_max <= 255 else (2 if x_max <= 65535 else 4)
        self.attr_bytes = 1 if attr_max <= 255 else 2
        self.p_bytes = 1 if periph_max <= 255 else 2
        if attr_max > 65535 or periph_max > 65535:
            raise ValueError("attributes above 65535 do not fit the compact wire format")
        off, o = {}, 0
        for name, nbytes in (("hdr", 16), ("gptr", 4 * (self.G + 1)), ("y", 4 * self.G), ("x", self.x_bytes * self.n_cap),
                             ("src", 4 * self.e_cap), ("dst", 4 * self.e_cap),
                             ("attr", self.attr_bytes * self.e_cap * self.K),
                             ("pea", self.p_bytes * self.n_cap * self.K * self.met * 2),
                             ("pca", self.p_bytes * self.n_cap * self.K * self.hp1)):
            off[name] = (o, nbytes)
            o += _align(nbytes)
        self.offsets, self.nbytes = off, o

    _DT = {1: np.uint8, 2: np.uint16, 4: np.int32}

    def host_buffer(self):
        return torch.zeros(self.nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())

    def _view(self, flat_np, name, dtype):
        o, n = self.offsets[name]
        return flat_np[o:o + n].view(dtype)

    def pack(self, b, flat):
        """b: a collated batch in the reference's layout on the HOST (kpgnn_b200.model.Batch of CPU tensors);
        flat: a `host_buffer()`.  Raises when the batch exceeds a capacity or an element width."""
        f = flat.numpy()
        N, E = int(b.x.size(0)), int(b.edge_index.size(1))
        if N > self.n_cap or E > self.e_cap or b.num_graphs != self.G:
            raise ValueError("batch (N=%d, E=%d, graphs=%d) exceeds the wire capacity (N=%d, E=%d, graphs=%d)"
                             % (N, E, b.num_graphs, self.n_cap, self.e_cap, self.G))
        self._view(f, "hdr", np.int32)[:4] = (N, E, self.G, 0)
        batch = b.batch.numpy()
        gptr = np.searchsorted(batch, np.arange(self.G + 1))
        self._view(f, "gptr", np.int32)[:] = gptr
        self._view(f, "y", np.float32)[:] = b.y.numpy().reshape(-1)

        def put(name, t, width, n):
            a = t.numpy().reshape(-1)
            if a.size and (a.min() < 0 or a.max() > (1 << (8 * width)) - 1 - (width == 4)):
                raise ValueError("%s holds values outside its %d-byte wire width" % (name, width))
            self._view(f, name, self._DT[width])[:n] = a
        put("x", b.x, self.x_bytes, N)
        put("src", b.edge_index[0], 4, E)
        put("dst", b.edge_index[1], 4, E)
        put("attr", b.edge_attr, self.attr_bytes, E * self.K)
        if b.pe_attr is not None and int(b.pe_attr.abs().sum()) != 0:
            raise ValueError("the compact wire format drops pe_attr (identically zero in the reference's extractor)")
        put("pea", b.peripheral_edge_attr, self.p_bytes, N * self.K * self.met * 2)
        put("pca", b.peripheral_configuration_attr, self.p_bytes, N * self.K * self.hp1)
        return flat


class DeviceWire(object):
    """Static device tensors of one batch stream: `stage` (the flat compact buffer, target of the H2D copy) and the
    int64 wire tensors at capacity that `unpack()` fills; `batch` is the kpgnn_b200.model.Batch over them."""

    def __init__(self, spec, device):
        self.spec, self.device = spec, torch.device(device)
        s, dev = spec, self.device
        self.stage = torch.zeros(s.nbytes, dtype=torch.uint8, device=dev)
        self.n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        i64 = dict(dtype=torch.int64, device=dev)
        self.batch = Batch(num_graphs=s.G, num_nodes=s.n_cap, n_dev=self.n_dev,
                           x=torch.zeros(s.n_cap, **i64), edge_index=torch.zeros((2, s.e_cap), **i64),
                           edge_attr=torch.zeros((s.e_cap, s.K), **i64),
                           pe_attr=torch.zeros((s.n_cap, s.K - 1), **i64) if s.K > 1 else None,
                           peripheral_edge_attr=torch.zeros((s.n_cap, s.K, s.met, 2), **i64),
                           peripheral_configuration_attr=torch.zeros((s.n_cap, s.K, s.hp1), **i64),
                           batch=torch.zeros(s.n_cap, **i64), y=torch.zeros(s.G, dtype=torch.float32, device=dev))
        d = _lib.WireDesc()
        d.n_cap, d.e_cap, d.g, d.K, d.met, d.hp1 = s.n_cap, s.e_cap, s.G, s.K, s.met, s.hp1
        d.x_bytes, d.attr_bytes, d.p_bytes = s.x_bytes, s.attr_bytes, s.p_bytes
        base = self.stage.data_ptr()
        for name in ("hdr", "gptr", "x", "src", "dst", "attr", "pea", "pca"):
            setattr(d, name, base + s.offsets[name][0])
        b = self.batch
        d.o_x, d.o_batch, d.o_ei, d.o_ea = b.x.data_ptr(), b.batch.data_ptr(), b.edge_index.data_ptr(), b.edge_attr.data_ptr()
        d.o_pea, d.o_pca = b.peripheral_edge_attr.data_ptr(), b.peripheral_configuration_attr.data_ptr()
        d.o_n = self.n_dev.data_ptr()
        self.desc = d
        o, n = s.offsets["y"]
        self._y_src = self.stage[o:o + n].view(torch.float32)

    def unpack(self):
        """stage -> wire tensors (one kernel + the 512-byte target copy), on the current stream; capturable."""
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib().kp_wire_unpack(C.byref(self.desc), st), "kp_wire_unpack")
        self.batch.y.copy_(self._y_src)
        # The wire tensors were rewritten by a raw kernel: bump their version counters so that the caches keyed on them
        # (graph plan, peripheral index) refresh on the next eager use.  Under stream capture nothing may re-plan
        # implicitly -- the capturing caller refreshes the plan explicitly (kpgnn_b200/train.py).
        if not torch.cuda.is_current_stream_capturing():
            ts = [t for t in (self.batch.edge_index, self.batch.edge_attr, self.batch.peripheral_edge_attr,
                              self.batch.peripheral_configuration_attr, self.batch.x, self.batch.batch)]
            bump = getattr(torch._C._autograd, "_unsafe_set_version_counter", None)
            if bump is not None:
                bump(ts, [t._version + 1 for t in ts])
            else:                       # older / newer torch without the hook: an in-place no-op bumps the counter too
                for t in ts:
                    t.add_(0)
        return self.batch
