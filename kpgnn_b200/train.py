"""Training-step driver for the K-hop path on one GPU of a data-parallel job: the
caller-side loop of the reference's train scripts (train_ZINC.py:29-47: for every batch `.to(device)`, forward, L1 loss,
backward, Adam) restated for static buffers, so that ONE captured CUDA graph serves a stream of batches of different
sizes:

  host:    WireSpec.pack(batch) -> flat pinned buffer (compact wire format, kpgnn_b200/wire.py)          [loader side]
  copy:    flat -> staging buffer on a copy stream while the previous step computes
  graph:   staging -> step inputs (D2D) | kp_wire_unpack | plan refresh (side stream) | peripheral index |
           forward | backward | [gradients packed for the all-reduce] | Adam
  world>1: forward+backward graph, NCCL all-reduce of the flat gradient, Adam graph

Capacities (nodes, K-hop edges, plan entries, largest hop attribute) are fixed when the trainer is built -- a loader
knows them from preprocessing -- and every refresh is validated against them at the step's next sync point.
"""
import os

import torch

from . import plan as kplan
from .dist import FlatGradients, PeerGradients
from .encoders import peripheral_index
from .model import l1_loss
from .optim import FusedAdam
from .wire import DeviceWire, WireSpec


class Bounds(object):
    """What the stream's largest batch needs: plan entries and the largest hop-1 / hop-k attribute values."""

    def __init__(self, nnz_cap, max_attr0, max_attrk):
        self.nnz_cap, self.max_attr0, self.max_attrk = int(nnz_cap), int(max_attr0), int(max_attrk)


def fit_spec(host_batches, K, max_edge_type, max_hop_num, headroom=1.08):
    """WireSpec + Bounds covering `host_batches` (kpgnn_b200.model.Batch objects of CPU tensors), with headroom on the
    node / edge / entry capacities for batches the sample did not contain."""
    n = max(int(b.x.size(0)) for b in host_batches)
    e = max(int(b.edge_index.size(1)) for b in host_batches)
    nnz = max(int((b.edge_attr != 0).sum()) for b in host_batches)
    a0 = max(int(b.edge_attr[:, 0].max()) for b in host_batches)
    ak = max(int(b.edge_attr[:, 1:].max()) if b.edge_attr.size(1) > 1 else 0 for b in host_batches)
    pm = max(max(int(b.peripheral_edge_attr.max()), int(b.peripheral_configuration_attr.max())) for b in host_batches)
    xm = max(int(b.x.max()) for b in host_batches)
    g = host_batches[0].num_graphs

    def up(v):
        return (int(v * headroom) + 63) // 64 * 64
    spec = WireSpec(up(n), up(e), g, K, max_edge_type, max_hop_num, attr_max=max(a0, ak), periph_max=pm, x_max=xm)
    return spec, Bounds(up(nnz), a0, ak)


class Trainer(object):
    """model: a graph-level model over kpgnn_b200.model.KPGNNPlusBackbone (fused layer stack: padded capacity batches
    need it); spec / bounds: see fit_spec()."""

    def __init__(self, model, spec, bounds, device, world=1, lr=1e-3, loss_fn=l1_loss, use_graph=True, adam_eps=1e-8):
        self.model, self.spec, self.bounds = model, spec, bounds
        self.device, self.world, self.loss_fn, self.use_graph = torch.device(device), world, loss_fn, use_graph
        self.wire = DeviceWire(spec, device)
        self.dev = self.wire.batch
        self.stage_next = torch.zeros(spec.nbytes, dtype=torch.uint8, device=device)
        self.copy_stream = torch.cuda.Stream(device)
        self.plan_stream = torch.cuda.Stream(device)
        self.staged = self.consumed = None
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.grads = self._make_grads() if world > 1 else None
        self.opt = FusedAdam(self.params, lr=lr, eps=adam_eps)
        self.loss = None
        self.graph = self.graph_opt = None
        self.launches_per_step = 0
        self.idx_buf = peripheral_index(self.dev.peripheral_edge_attr, self.dev.peripheral_configuration_attr)
        self.plan_obj = None
        self.fuse_head = os.environ.get("KP_FUSED_HEAD", "1") != "0"
        self._loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._e2e_i, self._e2e_prev = 0, None

    def _make_grads(self):
        """world > 1: the library's peer-memory exchange (csrc/peer.cu) when every rank can open every other rank's
        gradient block; otherwise (KP_PEER_ALLREDUCE=0, or no P2P path) the process group's all-reduce."""
        import torch.distributed as dist
        want = os.environ.get("KP_PEER_ALLREDUCE", "1") != "0"
        grads, err = None, None
        if want:
            try:
                grads = PeerGradients(self.params, barrier=False)    # the agreement below is the barrier
            except Exception as e:                      # noqa: BLE001 -- any set-up failure selects the NCCL exchange
                err = e
        ok = torch.tensor([1 if grads is not None else 0], device=self.device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # all ranks or none
        if int(ok.item()) == 1:
            return grads
        if want and dist.get_rank() == 0:
            import warnings
            warnings.warn("peer-memory gradient exchange unavailable (%s); using the process group's all-reduce" % (err,))
        return FlatGradients(self.params)

    # ---- derived per-batch state, recomputed every step into static buffers
    def _tag_idx(self):
        d = self.dev
        d._peripheral_idx = ((d.peripheral_edge_attr._version, d.peripheral_configuration_attr._version), self.idx_buf)

    def _plan(self):
        if self.plan_obj is None:
            kplan.reserve_capacity(self.bounds.nnz_cap)
            try:
                p, _ = kplan.get_plan(self.dev.edge_index, self.dev.edge_attr, self.dev.x.size(0))
            finally:
                kplan.reserve_capacity(0)
            # validated bounds of the whole stream, not of the first batch (they select kernels and guard the tables)
            p.max_attr0, p.max_attrk = max(p.max_attr0, self.bounds.max_attr0), max(p.max_attrk, self.bounds.max_attrk)
            self.plan_obj = p
        return self.plan_obj

    def refresh_derived(self):
        p = self._plan()
        base = self.dev.edge_attr
        ok = kplan.refresh_plan_async(p, self.dev.edge_index, base, base.size(1), self.plan_stream)
        assert ok
        kplan.mark_current(self.dev.edge_index, self.dev.edge_attr, self.dev.x.size(0))
        self.idx_buf.copy_(peripheral_index(self.dev.peripheral_edge_attr, self.dev.peripheral_configuration_attr))
        self._tag_idx()

    # ---- one optimisation step on whatever the staging buffer holds
    def _zero(self):
        for p in self.params:
            p.grad = None

    def _fwd_bwd(self):
        self.wire.unpack()
        self.refresh_derived()
        self._zero()
        loss = None
        if self.loss_fn is l1_loss and self.fuse_head and hasattr(self.model, "fused_loss"):
            loss = self.model.fused_loss(self.dev, "l1")          # head + loss as one kernel each way (kp_head_*)
        if loss is None:
            loss = self.loss_fn(self.model(self.dev), self.dev.y)
        loss.backward()
        if self.plan_obj is not None and self.plan_obj.tail is not None:     # late join of the plan stream (host statistics)
            torch.cuda.current_stream(self.device).wait_event(self.plan_obj.tail)
            self.plan_obj.tail = None
        if self.grads is not None:
            self.grads.gather_()
        return loss.detach()

    def _step_eager(self):
        loss = self._fwd_bwd()
        if self.grads is not None:
            self.grads.allreduce_mean_(self.world)
        self.opt.step()
        return loss

    def load(self, host_flat):
        """Synchronously place one packed batch in the staging buffer (set-up, tests)."""
        self.wire.stage.copy_(host_flat, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def capture(self, host_flat):
        """Warm up on `host_flat` (allocations, plan capacity), then capture the step."""
        from . import _lib
        kplan.deferred_checks(True)
        self.load(host_flat)
        if not self.use_graph:
            n0 = _lib.launch_count()
            self.loss = self._step_eager()
            self.launches_per_step = _lib.launch_count() - n0
            return
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(3):
                self._step_eager()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        if self.world == 1:
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd()
                self.opt.step()
        else:
            # world > 1: the NCCL all-reduce of the flat gradient is captured INSIDE the step graph (one launch per
            # step, no host round trip between backward, collective and Adam).  If this NCCL / driver combination
            # refuses to capture a collective, fall back to: forward+backward graph, eager all-reduce, Adam graph.
            in_graph = isinstance(self.grads, PeerGradients) or os.environ.get("KP_NCCL_IN_GRAPH", "0") == "1"
            try:
                if not in_graph:
                    raise RuntimeError("collective outside the graph")
                with torch.cuda.graph(self.graph):
                    self.loss = self._fwd_bwd()
                    self.grads.allreduce_mean_(self.world)
                    self.opt.step()
                self.graph_opt = None
            except Exception:
                torch.cuda.synchronize(self.device)
                self.graph = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(self.graph):
                    self.loss = self._fwd_bwd()
                self.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_opt):
                    self.opt.step()
        self.launches_per_step = _lib.launch_count() - n0
        torch.cuda.synchronize(self.device)

    def replay(self):
        if self.graph is None:
            self.loss = self._step_eager()
            return
        self.graph.replay()
        if self.graph_opt is not None:
            self.grads.allreduce_mean_(self.world)
            self.graph_opt.replay()

    # ---- input pipeline
    def prefetch(self, host_flat):
        """Host -> device copy of the NEXT batch on the copy stream (pinned memory, one transfer)."""
        cs = self.copy_stream
        if self.consumed is not None:
            cs.wait_event(self.consumed)
        with torch.cuda.stream(cs):
            self.stage_next.copy_(host_flat, non_blocking=True)
            self.staged = torch.cuda.Event()
            self.staged.record(cs)

    def hand_over(self):
        st = torch.cuda.current_stream(self.device)
        st.wait_event(self.staged)
        self.wire.stage.copy_(self.stage_next, non_blocking=True)
        self.consumed = torch.cuda.Event()
        self.consumed.record(st)

    def step_resident(self):
        """The staged batch is already in HBM: unpack + plan + step."""
        self.replay()

    def step_e2e(self, next_host_flat):
        """One step with HOST inputs: the batch staged by the previous call is handed to the step, the step is launched,
        the next batch's upload is issued behind it, and the loss is read back (train_ZINC.py:45)."""
        self.hand_over()
        self.replay()
        self.prefetch(next_host_flat)
        val = self.loss.item()
        self.validate()
        return val

    def step_e2e_pipelined(self, next_host_flat):
        """step_e2e without a host stall per step: the loss of step i is copied to pinned memory behind step i and READ
        while step i+1 runs (one value per step, one step late -- the accumulation of train_ZINC.py:45 is unchanged); the
        deferred plan checks read the sticky statistics the same way.  Returns the previous step's loss (None on the
        first call); drain() returns the last one."""
        self.hand_over()
        self.replay()
        st = torch.cuda.current_stream(self.device)
        slot = self._e2e_i & 1
        self._loss_host[slot].copy_(self.loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(st)
        self.prefetch(next_host_flat)
        prev, self._e2e_prev = self._e2e_prev, (ev, slot)
        self._e2e_i += 1
        if prev is None:
            return None
        prev[0].synchronize()
        val = float(self._loss_host[prev[1]])
        self._validate_lagged()
        return val

    def drain(self):
        """Loss of the last pipelined step (waits for it) + a full validation."""
        prev, self._e2e_prev = self._e2e_prev, None
        if prev is None:
            return None
        prev[0].synchronize()
        val = float(self._loss_host[prev[1]])
        self.validate()
        return val

    def _validate_lagged(self):
        if self.plan_obj is not None:
            self.plan_obj.validate_lagged()
        if isinstance(self.grads, PeerGradients):
            self._validations = getattr(self, "_validations", 0) + 1
            if self._validations % 64 == 1:
                self.grads.check()

    def validate(self):
        if self.plan_obj is not None:
            self.plan_obj.validate()
        if isinstance(self.grads, PeerGradients):
            self._validations = getattr(self, "_validations", 0) + 1
            if self._validations % 16 == 1:
                self.grads.check()


class GraphedStep(object):
    """forward + backward + Adam of ANY model built on the drop-in layers (the native backbones, or the reference's own
    unmodified models/GNNs.py after install_dropin()) on a FIXED-SHAPE batch as one CUDA graph: the eager step of these
    models is launch-bound (hundreds of 2-10 us kernels), a graph replay is not.  `batch` and `y` are the static inputs:
    copy a new batch of the same shapes INTO their tensors between replays (the plan cache follows the tensors' version
    counters outside capture only, so contents that change the K-hop structure need kpgnn_b200.plan.refresh_plan +
    mark_current, as Trainer does; identical structure -- e.g. new node features / targets -- needs nothing)."""

    def __init__(self, model, batch, y, loss_fn=l1_loss, lr=1e-3, warmup=3):
        self.model, self.batch, self.y, self.loss_fn = model, batch, y, loss_fn
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = FusedAdam(self.params, lr=lr)
        dev = self.params[0].device
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()

    def _step(self):
        for p in self.params:
            p.grad = None
        loss = self.loss_fn(self.model(self.batch), self.y)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def replay(self):
        self.graph.replay()
        return self.loss
