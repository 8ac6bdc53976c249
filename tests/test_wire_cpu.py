"""Host side of the compact wire format (kpgnn_b200/wire.py WireSpec.pack, kpgnn_b200/train.py fit_spec) and of the
chunk planner for the node-range backward (kpgnn_b200/ops.py backward_chunks): pure host logic, no GPU.  The device side
(kp_wire_unpack) is checked against the same layout in tests/test_varbatch_gpu.py."""
import numpy as np
import pytest
import torch

from kpgnn_b200 import synth
from tests.util import collate

ARGS = (4, 50, 6, 3, 50, 50, "spd")


def _batch(num_graphs, seed):
    from kpgnn_b200.model import Batch
    return Batch(**collate(synth.zinc_like_graphs(num_graphs, seed=seed), ARGS))


def _decode(spec, flat):
    """Independent reading of the documented layout: offsets from the spec, little-endian fields of the given widths."""
    f = flat.numpy()
    widths = {1: np.uint8, 2: np.uint16, 4: np.int32}

    def view(name, dt):
        o, n = spec.offsets[name]
        return f[o:o + n].view(dt)
    N, E, G, _ = view("hdr", np.int32)[:4]
    return {"N": int(N), "E": int(E), "G": int(G), "gptr": view("gptr", np.int32).copy(), "y": view("y", np.float32).copy(),
            "x": view("x", widths[spec.x_bytes])[:N].astype(np.int64),
            "src": view("src", np.int32)[:E].astype(np.int64), "dst": view("dst", np.int32)[:E].astype(np.int64),
            "attr": view("attr", widths[spec.attr_bytes])[:E * spec.K].astype(np.int64).reshape(E, spec.K),
            "pea": view("pea", widths[spec.p_bytes])[:N * spec.K * spec.met * 2].astype(np.int64),
            "pca": view("pca", widths[spec.p_bytes])[:N * spec.K * spec.hp1].astype(np.int64)}


def test_pack_layout_roundtrip_and_capacities():
    from kpgnn_b200.train import fit_spec
    hbs = [_batch(10, s) for s in (1, 2, 3)]
    spec, bounds = fit_spec(hbs, 4, 3, 6)
    assert spec.n_cap >= max(int(b.x.size(0)) for b in hbs) and spec.e_cap >= max(int(b.edge_index.size(1)) for b in hbs)
    for b in hbs:
        d = _decode(spec, spec.pack(b, spec.host_buffer()))
        N, E = int(b.x.size(0)), int(b.edge_index.size(1))
        assert (d["N"], d["E"], d["G"]) == (N, E, 10)
        assert np.array_equal(d["x"], b.x.numpy().reshape(-1))
        assert np.array_equal(d["src"], b.edge_index[0].numpy()) and np.array_equal(d["dst"], b.edge_index[1].numpy())
        assert np.array_equal(d["attr"], b.edge_attr.numpy())
        assert np.array_equal(d["pea"], b.peripheral_edge_attr.numpy().reshape(-1))
        assert np.array_equal(d["pca"], b.peripheral_configuration_attr.numpy().reshape(-1))
        assert np.allclose(d["y"], b.y.numpy().reshape(-1))
        # graph pointer: node ranges of the graphs, consistent with the batch vector
        bt = b.batch.numpy()
        assert d["gptr"][0] == 0 and d["gptr"][-1] == N
        for g in range(10):
            assert np.all(bt[d["gptr"][g]:d["gptr"][g + 1]] == g)
    # the flat buffer is much smaller than the int64 wire layout it replaces
    assert spec.nbytes < 0.4 * hbs[0].nbytes()


def test_pack_refuses_what_does_not_fit():
    from kpgnn_b200.wire import WireSpec
    b = _batch(6, 7)
    N, E = int(b.x.size(0)), int(b.edge_index.size(1))
    ok = WireSpec(N, E, 6, 4, 3, 6)
    ok.pack(b, ok.host_buffer())
    with pytest.raises(ValueError):
        s = WireSpec(N - 1, E, 6, 4, 3, 6)
        s.pack(b, s.host_buffer())
    with pytest.raises(ValueError):
        s = WireSpec(N, E - 1, 6, 4, 3, 6)
        s.pack(b, s.host_buffer())
    with pytest.raises(ValueError):
        s = WireSpec(N, E, 7, 4, 3, 6)
        s.pack(b, s.host_buffer())
    big = _batch(6, 7)
    big.edge_attr = big.edge_attr.clone()
    big.edge_attr[0, 1] = 300                                   # does not fit one byte
    with pytest.raises(ValueError):
        ok.pack(big, ok.host_buffer())
    wide = WireSpec(N, E, 6, 4, 3, 6, attr_max=300)            # two-byte attributes
    d = _decode(wide, wide.pack(big, wide.host_buffer()))
    assert d["attr"][0, 1] == 300
    with pytest.raises(ValueError):
        WireSpec(N, E, 6, 4, 3, 6, attr_max=70000)


class _FakePlan(object):
    def __init__(self, sizes):
        self.N = int(sum(sizes))
        self._bp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)

    def block_ptr_host(self):
        return self._bp


def test_backward_chunks_follow_graph_boundaries():
    from kpgnn_b200 import ops
    rng = np.random.default_rng(0)
    sizes = rng.integers(5, 40, size=200)
    plan = _FakePlan(sizes)
    old = (ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES)
    try:
        ops.CHUNK_BWD_MIN_BYTES = None
        assert ops.backward_chunks(plan, 8, 104) is None                      # off by default
        ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES = 0, 300 * 4 * 8 * 104
        ch = ops.backward_chunks(plan, 8, 104)
        ends = set(np.cumsum(sizes).tolist()) | {0}
        assert ch[0] == 0 and ch[-1] == plan.N and all(c in ends for c in ch) and ch == sorted(set(ch))
        assert max(b - a for a, b in zip(ch[:-1], ch[1:])) <= 300                # every chunk within the target
        ops.CHUNK_BWD_GS_BYTES = 10 * 4 * 8 * 104                              # target below one graph: one graph per chunk
        ch = ops.backward_chunks(_FakePlan([50, 60, 70]), 8, 104)
        assert ch == [0, 50, 110, 180]
        ops.CHUNK_BWD_MIN_BYTES = 1 << 40                                      # batch below the threshold: single call
        assert ops.backward_chunks(plan, 8, 104) is None
    finally:
        ops.CHUNK_BWD_MIN_BYTES, ops.CHUNK_BWD_GS_BYTES = old
