"""ctypes binding of libkpgnn_b200.so (C ABI declared in include/kpgnn.h).

There is NO fallback: if the shared library is missing or fails to load, importing this module raises, and so
does every product entry point.  The oracle under `oracle/` is never reachable from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KPGNN_B200_LIB points at an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("KPGNN_B200_LIB") or os.path.join(_HERE, "libkpgnn_b200.so")

ABI_VERSION = 17


class KpError(RuntimeError):
    pass


class PlanInput(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("attr", C.c_void_p), ("attr_stride", C.c_int64),
                ("N", C.c_int32), ("E", C.c_int32), ("K", C.c_int32), ("self_loops", C.c_int32)]


class AggDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("Kplan", C.c_int32), ("k", C.c_int32), ("d", C.c_int32),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("attr16", C.c_void_p),
                ("rowptrT", C.c_void_p), ("colT", C.c_void_p), ("dinv", C.c_void_p), ("indeg", C.c_void_p),
                ("X", C.c_void_p), ("x_node_stride", C.c_int64), ("x_hop_stride", C.c_int64),
                ("P", C.c_void_p), ("p_node_stride", C.c_int64), ("p_hop_stride", C.c_int64),
                ("T0", C.c_void_p), ("Tk", C.c_void_p), ("rows0", C.c_int32), ("rowsk", C.c_int32),
                ("theta", C.c_void_p), ("eps", C.c_void_p), ("act", C.c_int32), ("fuse", C.c_int32),
                ("amax0", C.c_int32), ("amaxk", C.c_int32),
                ("dx_node_stride", C.c_int64), ("dx_hop_stride", C.c_int64),
                ("dx_accumulate", C.c_int32), ("node_base", C.c_int32),
                ("geo_alphas", C.c_void_p), ("geo_dalphas", C.c_void_p), ("leaf_stream", C.c_void_p),
                ("block_ptr", C.c_void_p), ("block_stats", C.c_void_p), ("num_blocks", C.c_int32),
                ("max_block_nodes", C.c_int32)]


class ExtractInput(C.Structure):
    _fields_ = [("G", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("n_max", C.c_int32),
                ("gptr", C.c_void_p), ("node_graph", C.c_void_p), ("pair_off", C.c_void_p),
                ("erow", C.c_void_p), ("ecol", C.c_void_p), ("emult", C.c_void_p), ("etype", C.c_void_p),
                ("kernel", C.c_int32), ("cap", C.c_int32),
                ("max_edge_attr_num", C.c_int32), ("max_hop_num", C.c_int32), ("max_edge_type", C.c_int32),
                ("max_edge_count", C.c_int32), ("max_distance_count", C.c_int32), ("max_type_value", C.c_int32)]


class TsumDesc(C.Structure):
    _fields_ = [("R", C.c_int32), ("S", C.c_int32), ("d", C.c_int32), ("table_rows", C.c_int32),
                ("idx", C.c_void_p), ("slot_off", C.c_int32 * 32), ("num_ranges", C.c_int32),
                ("range_slot", C.c_int32 * 17), ("range_row", C.c_int32 * 17)]


class DenseDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("Cin", C.c_int32), ("Cout", C.c_int32),
                ("X", C.c_void_p),
                ("W1", C.c_void_p), ("b1", C.c_void_p), ("g1", C.c_void_p), ("be1", C.c_void_p),
                ("W2", C.c_void_p), ("b2", C.c_void_p), ("g2", C.c_void_p), ("be2", C.c_void_p),
                ("g3", C.c_void_p), ("be3", C.c_void_p), ("R", C.c_void_p),
                ("eps1", C.c_float), ("eps2", C.c_float), ("eps3", C.c_float),
                ("mom1", C.c_float), ("mom2", C.c_float), ("mom3", C.c_float),
                ("rm1", C.c_void_p), ("rv1", C.c_void_p), ("rm2", C.c_void_p), ("rv2", C.c_void_p),
                ("rm3", C.c_void_p), ("rv3", C.c_void_p),
                ("nbt1", C.c_void_p), ("nbt2", C.c_void_p), ("nbt3", C.c_void_p),
                ("Y1", C.c_void_p), ("Y2", C.c_void_p), ("Z2", C.c_void_p), ("stats", C.c_void_p),
                ("out_stride", C.c_int64), ("r_stride", C.c_int64), ("dout_stride", C.c_int64),
                ("dr_stride", C.c_int64), ("dR", C.c_void_p), ("barrier", C.c_void_p),
                ("leaf_stream", C.c_void_p), ("n_dev", C.c_void_p)]


class AttnDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("K", C.c_int32), ("d", C.c_int32), ("pad", C.c_int32),
                ("x", C.c_void_p), ("x_node_stride", C.c_int64), ("x_hop_stride", C.c_int64),
                ("w_ih", C.c_void_p * 2), ("w_hh", C.c_void_p * 2), ("b_ih", C.c_void_p * 2), ("b_hh", C.c_void_p * 2)]


class FoldDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("H_in", C.c_int32), ("H_out", C.c_int32), ("gate_act", C.c_int32),
                ("E", C.c_void_p * 16), ("W", C.c_void_p * 16), ("w_stride", C.c_int64 * 16), ("rows", C.c_int32 * 16),
                ("gate", C.c_int32 * 16), ("row_off", C.c_int32 * 17), ("pad", C.c_int32),
                ("gate_raw", C.c_void_p * 2), ("bias", C.c_void_p * 2), ("bias_mult", C.c_float * 2)]


class FoldGrads(C.Structure):
    _fields_ = [("dE", C.c_void_p * 16), ("dW", C.c_void_p * 16), ("dbias", C.c_void_p * 2),
                ("dgate_raw", C.c_void_p * 2)]


PEER_MAX, PEER_CTAS, PEER_HANDLE_BYTES = 8, 64, 64
PEER_FLAG_BYTES = 2 * PEER_CTAS * PEER_MAX * 4


class PeerDesc(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n", C.c_int64), ("block", C.c_void_p * PEER_MAX),
                ("out", C.c_void_p), ("epoch", C.c_void_p), ("error", C.c_void_p), ("scale", C.c_float)]


class HeadDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("G", C.c_int32), ("mean", C.c_int32), ("loss_kind", C.c_int32),
                ("pad", C.c_int32), ("rep", C.c_void_p), ("rep_stride", C.c_int64), ("rep_stride_out", C.c_int64),
                ("batch", C.c_void_p), ("n_dev", C.c_void_p), ("w", C.c_void_p), ("b", C.c_void_p), ("y", C.c_void_p)]


class WireDesc(C.Structure):
    _fields_ = [("n_cap", C.c_int32), ("e_cap", C.c_int32), ("g", C.c_int32), ("K", C.c_int32), ("met", C.c_int32),
                ("hp1", C.c_int32), ("x_bytes", C.c_int32), ("attr_bytes", C.c_int32), ("p_bytes", C.c_int32),
                ("pad", C.c_int32), ("hdr", C.c_void_p), ("gptr", C.c_void_p), ("x", C.c_void_p), ("src", C.c_void_p),
                ("dst", C.c_void_p), ("attr", C.c_void_p), ("pea", C.c_void_p), ("pca", C.c_void_p),
                ("o_x", C.c_void_p), ("o_batch", C.c_void_p), ("o_ei", C.c_void_p), ("o_ea", C.c_void_p),
                ("o_pea", C.c_void_p), ("o_pca", C.c_void_p), ("o_n", C.c_void_p)]


class ThetaBatch(C.Structure):
    _fields_ = [("L", C.c_int32), ("d", C.c_int32), ("alphas", C.c_void_p * 32), ("theta", C.c_void_p * 32),
                ("k", C.c_int32 * 32)]


class PgradDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("K", C.c_int32), ("d", C.c_int32), ("L", C.c_int32),
                ("dagg", C.c_void_p * 32), ("theta", C.c_void_p * 32), ("k", C.c_int32 * 32)]


ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2

# name -> (restype, argtypes); must list every symbol include/kpgnn.h declares (tests/test_abi.py checks it)
_SIGNATURES = {
    "kp_last_error": (C.c_char_p, []),
    "kp_abi_version": (C.c_int, []),
    "kp_launch_count": (C.c_uint64, []),
    "kp_plan_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "kp_plan_count": (C.c_int, [C.POINTER(PlanInput), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_plan_fill": (C.c_int, [C.POINTER(PlanInput), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_plan_clamp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kp_plan_blocks_workspace_bytes": (C.c_int, [C.c_int32, C.POINTER(C.c_size_t)]),
    "kp_plan_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_agg_forward": (C.c_int, [C.POINTER(AggDesc), C.c_void_p, C.c_void_p]),
    "kp_agg_set_force_generic": (C.c_int, [C.c_int]),
    "kp_agg_set_launch_geometry": (C.c_int, [C.c_int, C.c_int]),
    "kp_agg_backward_workspace_bytes": (C.c_int, [C.POINTER(AggDesc), C.POINTER(C.c_size_t)]),
    "kp_agg_backward_chunkable": (C.c_int, [C.POINTER(AggDesc), C.POINTER(C.c_int32)]),
    "kp_agg_backward": (C.c_int, [C.POINTER(AggDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_table_sum_forward": (C.c_int, [C.POINTER(TsumDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "kp_table_sum_set_smem_cap": (C.c_int, [C.c_size_t]),
    "kp_table_sum_backward_workspace_bytes": (C.c_int, [C.POINTER(TsumDesc), C.POINTER(C.c_size_t)]),
    "kp_table_sum_backward": (C.c_int, [C.POINTER(TsumDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "kp_bn_max_rows": (C.c_int, []),
    "kp_bn_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "kp_bn_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kp_dense_block_max_rows": (C.c_int, [C.c_int32, C.c_int32]),
    "kp_dense_block_set_mma": (C.c_int, [C.c_int]),
    "kp_dense_block_workspace_bytes": (C.c_int, [C.POINTER(DenseDesc), C.POINTER(C.c_size_t),
                                                 C.POINTER(C.c_size_t)]),
    "kp_dense_block_forward": (C.c_int, [C.POINTER(DenseDesc), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_dense_block_backward": (C.c_int, [C.POINTER(DenseDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_geometric_theta_forward_batched": (C.c_int, [C.POINTER(ThetaBatch), C.c_void_p]),
    "kp_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_double, C.c_double, C.c_float,
                               C.c_void_p, C.c_void_p]),
    "kp_attn_combine_forward": (C.c_int, [C.POINTER(AttnDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "kp_attn_combine_backward_workspace_bytes": (C.c_int, [C.POINTER(AttnDesc), C.POINTER(C.c_size_t)]),
    "kp_attn_combine_backward": (C.c_int, [C.POINTER(AttnDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_void_p]),
    "kp_fold_forward": (C.c_int, [C.POINTER(FoldDesc), C.c_void_p, C.c_void_p]),
    "kp_fold_backward": (C.c_int, [C.POINTER(FoldDesc), C.c_void_p, C.POINTER(FoldGrads), C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "kp_wire_unpack": (C.c_int, [C.POINTER(WireDesc), C.c_void_p]),
    "kp_peer_block_bytes": (C.c_size_t, [C.c_int64]),
    "kp_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "kp_peer_free": (C.c_int, [C.c_void_p]),
    "kp_peer_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kp_peer_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "kp_peer_release": (C.c_int, [C.c_void_p]),
    "kp_peer_allreduce_mean": (C.c_int, [C.POINTER(PeerDesc), C.c_void_p]),
    "kp_head_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "kp_head_forward": (C.c_int, [C.POINTER(HeadDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "kp_head_backward": (C.c_int, [C.POINTER(HeadDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kp_segment_sum": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_void_p, C.c_void_p]),
    "kp_peripheral_grad": (C.c_int, [C.POINTER(PgradDesc), C.c_void_p, C.c_void_p]),
    "kp_geometric_theta_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "kp_geometric_theta_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                              C.c_void_p]),
    "kp_extract_workspace_bytes": (C.c_int, [C.POINTER(ExtractInput), C.c_int64, C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t)]),
    "kp_extract_hops": (C.c_int, [C.POINTER(ExtractInput), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "kp_extract_emit": (C.c_int, [C.POINTER(ExtractInput), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int64, C.c_void_p]),
    "kp_extract_peripheral": (C.c_int, [C.POINTER(ExtractInput), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
}

_lib = None


def lib():
    """Loads the shared library on first use; raises KpError (never falls back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise KpError("libkpgnn_b200.so not built: run `python -m kpgnn_b200.build` "
                          "(there is no CPU or PyTorch fallback for the K-hop path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.kp_abi_version() != ABI_VERSION:
            raise KpError("libkpgnn_b200.so ABI %d != expected %d; rebuild" % (handle.kp_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


_BARRIERS = {}


def barrier_state(device):
    """Persistent zero-initialised grid-barrier state for kp_dense_block_* : one per (device, stream), so two dense
    blocks running concurrently on different streams never share a counter.  During stream capture the key is the
    capturing stream; the captured graph owns that state for its lifetime (it is never freed)."""
    import torch
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(device).cuda_stream)
    t = _BARRIERS.get(key)
    if t is None:
        t = torch.zeros(64, dtype=torch.int32, device=device)
        _BARRIERS[key] = t
    return t


_FOLD_WS = {}


def fold_workspace(device):
    """KP_FOLD_WORKSPACE_BYTES zero-initialised bytes per (device, stream) for kp_fold_backward (partials + self-resetting counter)."""
    import torch
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(device).cuda_stream)
    t = _FOLD_WS.get(key)
    if t is None:
        t = torch.zeros(2048, dtype=torch.uint8, device=device)
        _FOLD_WS[key] = t
    return t


def check(rc, what):
    if rc != 0:
        raise KpError("%s failed (%d): %s" % (what, rc, lib().kp_last_error().decode()))


def launch_count():
    return int(lib().kp_launch_count())
