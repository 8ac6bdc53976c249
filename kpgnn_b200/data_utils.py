"""K-hop neighbourhood / peripheral-subgraph extraction on the GPU -- mirror of the reference's `data_utils.py`.

  extract_multi_hop_neighbors(data, K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count,
                              max_distance_count, kernel)
      same name, argument meaning, field names, dtypes, ordering and quirks as data_utils.py:20-107 (it mutates
      and returns `data`, including the differently named/shaped fields of the E=0 branch, :37-44); the work is
      done by the CUDA kernels behind kp_extract_* (include/kpgnn.h).  Drop-in for the `pre_transform` closures
      of the reference's train scripts (e.g. train_ZINC.py:191-194).
  extract_batch(graphs, args, device)
      the B200-first entry: a whole list of raw graphs -> ONE collated batch in the reference's wire layout
      (what `Batch.from_data_list` would build from per-graph results), resident on the device.

The host only re-packs the raw edge lists into a CSR (format conversion, numpy); there is no CPU implementation
of the extraction here and none is reachable from this module.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import Batch


def _raw(g):
    """Accepts the dict form used by kpgnn_b200.synth or any object with PyG `Data` attributes."""
    if isinstance(g, dict):
        return g["num_nodes"], np.asarray(g["edge_index"]), g.get("edge_attr"), g.get("x"), g.get("y", 0.0)
    ei = g.edge_index
    ea = getattr(g, "edge_attr", None) if "edge_attr" in g else None
    x = getattr(g, "x", None)
    return (g.num_nodes, ei.cpu().numpy(), None if ea is None else ea.cpu().numpy(),
            None if x is None else x.cpu().numpy(), 0.0)


def pack_csr(graphs):
    """Raw graphs -> batch CSR by source with duplicate (src,dst) pairs merged (multiplicity, summed type), the
    same merge the reference's COO->dense conversions perform (data_utils.py:52-53).  Vectorised over the batch: the
    only per-graph Python work is collecting the arrays."""
    raws = [(g["num_nodes"], g["edge_index"], g.get("edge_attr")) if isinstance(g, dict) else _raw(g)[:3] for g in graphs]
    G = len(raws)
    ns = np.fromiter((r[0] for r in raws), dtype=np.int64, count=G)
    eis = [np.asarray(r[1], dtype=np.int64).reshape(2, -1) for r in raws]
    ecount = np.fromiter((e.shape[1] for e in eis), dtype=np.int64, count=G)
    gptr = np.zeros(G + 1, dtype=np.int64)
    np.cumsum(ns, out=gptr[1:])
    N = int(gptr[-1])
    pair_off = np.zeros(G + 1, dtype=np.int64)
    np.cumsum(ns * ns, out=pair_off[1:])
    E = int(ecount.sum())
    if E:
        ei = np.concatenate(eis, axis=1)
        nper = np.repeat(ns, ecount)
        bad = (ei < 0) | (ei >= nper)
        if bad.any():
            gi = int(np.searchsorted(np.cumsum(ecount), int(np.nonzero(bad.any(axis=0))[0][0]), side="right"))
            raise IndexError("edge_index out of range for a graph with %d nodes" % int(ns[gi]))
        typs = []
        for r, c in zip(raws, ecount):
            if r[2] is None:
                if c:
                    typs.append(np.full(c, 2, dtype=np.int64))                 # data_utils.py:46-50
            else:
                ea = np.asarray(r[2], dtype=np.int64).reshape(-1)
                if ea.shape[0] != c:
                    raise ValueError("edge_attr must be one integer per edge (got %s)" % (np.shape(r[2]),))
                if c:
                    typs.append(ea)
        typ = np.concatenate(typs)
        if typ.min() < 0:
            raise ValueError("negative edge types are not supported (the reference's bincount rejects them too)")
        off = np.repeat(gptr[:-1], ecount)
        key = (ei[0] + off) * N + (ei[1] + off)
        order = np.argsort(key, kind="stable")
        ks = key[order]
        first = np.empty(E, dtype=bool)
        first[0] = True
        np.not_equal(ks[1:], ks[:-1], out=first[1:])
        starts = np.flatnonzero(first)
        uniq = ks[starts]
        mult = np.diff(np.append(starts, E))
        tsum = np.add.reduceat(typ[order], starts)
        usrc, udst = uniq // N, uniq % N
    else:
        usrc = udst = mult = tsum = np.zeros(0, dtype=np.int64)
    erow = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(usrc, minlength=N), out=erow[1:])
    if tsum.size and tsum.max() >= 2 ** 20:
        raise ValueError("edge type values above 2^20 are not supported")
    return {
        "G": G, "N": N, "n_max": int(ns.max()) if G else 0, "total_pairs": int(pair_off[-1]),
        "gptr": gptr.astype(np.int32), "node_graph": np.repeat(np.arange(G, dtype=np.int32), ns),
        "pair_off": pair_off, "erow": erow.astype(np.int32), "ecol": udst.astype(np.int32),
        "emult": mult.astype(np.int32), "etype": tsum.astype(np.int32),
        "max_type_value": int(tsum.max()) if tsum.size else 0,
    }


def _extract_device(csr, K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count, max_distance_count,
                    kernel, device):
    """Runs the three kernels; returns device tensors in the reference layout."""
    lib = _lib.lib()
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.KpError("extraction runs on a CUDA device only (no CPU fallback); got %s" % device)
    if kernel not in ("spd", "gd"):
        raise ValueError("kernel must be 'spd' or 'gd'")
    if not (0 <= max_edge_attr_num <= 65534):
        raise ValueError("max_edge_attr_num must be in [0, 65534]")
    N = csr["N"]
    dv = csr.get("_device")                    # upload_csr(): the packed arrays already resident on the device
    if dv is None or dv["gptr"].device != device:
        dv = _upload(csr, device)
    ein = _lib.ExtractInput()
    ein.G, ein.N, ein.K, ein.n_max = csr["G"], N, K, csr["n_max"]
    ein.gptr, ein.node_graph, ein.pair_off = dv["gptr"].data_ptr(), dv["node_graph"].data_ptr(), dv["pair_off"].data_ptr()
    ein.erow, ein.ecol, ein.emult, ein.etype = (dv["erow"].data_ptr(), dv["ecol"].data_ptr(), dv["emult"].data_ptr(),
                                                dv["etype"].data_ptr())
    ein.kernel = 0 if kernel == "spd" else 1
    ein.cap = max(int(max_edge_attr_num), 1)
    ein.max_edge_attr_num, ein.max_hop_num, ein.max_edge_type = int(max_edge_attr_num), int(max_hop_num), int(max_edge_type)
    ein.max_edge_count, ein.max_distance_count = int(min(max_edge_count, 2 ** 31 - 1)), int(min(max_distance_count, 2 ** 31 - 1))
    ein.max_type_value = csr["max_type_value"]
    hop_b, scr_b = C.c_size_t(0), C.c_size_t(0)
    _lib.check(lib.kp_extract_workspace_bytes(C.byref(ein), csr["total_pairs"], C.byref(hop_b), C.byref(scr_b)),
               "kp_extract_workspace_bytes")
    W = torch.empty(max(hop_b.value // 2, 1), dtype=torch.int16, device=device)
    scratch = torch.empty(max(scr_b.value, 1), dtype=torch.uint8, device=device)
    eptr = torch.empty(N + 1, dtype=torch.int32, device=device)
    st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    _lib.check(lib.kp_extract_hops(C.byref(ein), W.data_ptr(), eptr.data_ptr(), scratch.data_ptr(), scratch.numel(), st),
               "kp_extract_hops")
    EK = int(eptr[-1].item()) if N > 0 else 0                    # host sync: sizes the outputs
    edge_index = torch.empty((2, EK), dtype=torch.int64, device=device)
    edge_attr = torch.empty((EK, K), dtype=torch.int64, device=device)
    _lib.check(lib.kp_extract_emit(C.byref(ein), W.data_ptr(), eptr.data_ptr(), edge_index.data_ptr(),
                                   edge_attr.data_ptr(), EK, st), "kp_extract_emit")
    out = {"edge_index": edge_index, "edge_attr": edge_attr,
           "pe_attr": torch.zeros((N, K - 1), dtype=torch.int64, device=device) if K > 1 else None,   # :91-96
           "peripheral_edge_attr": None, "peripheral_configuration_attr": None, "eptr": eptr}
    if max_hop_num > 0 and max_edge_type > 0:                     # data_utils.py:141
        pea = torch.empty((N, K, max_edge_type, 2), dtype=torch.int64, device=device)
        pca = torch.empty((N, K, max_hop_num + 1), dtype=torch.int64, device=device)
        _lib.check(lib.kp_extract_peripheral(C.byref(ein), W.data_ptr(), pea.data_ptr(), pca.data_ptr(),
                                             scratch.data_ptr(), scratch.numel(), st), "kp_extract_peripheral")
        out["peripheral_edge_attr"], out["peripheral_configuration_attr"] = pea, pca
    return out


_CSR_KEYS = ("pair_off", "gptr", "node_graph", "erow", "ecol", "emult", "etype")
_STAGE = {}          # device -> [pinned uint8 staging tensor (grow-only), event of the last copy out of it]


def _upload(csr, device, extra=None):
    """The packed arrays (and `extra` host arrays, e.g. node features) -> ONE pinned staging buffer -> one host-to-device
    copy; returns {name: device tensor view}.  Seven pageable copies (each a host-synchronous cudaMemcpy) were most of
    the extraction's end-to-end time at molecule batch sizes."""
    device = torch.device(device)
    items = [(k, np.ascontiguousarray(csr[k])) for k in _CSR_KEYS] + \
            [(k, np.ascontiguousarray(v)) for k, v in (extra or {}).items()]
    offs, total = [], 0
    for _, a in items:
        offs.append(total)
        total += (a.nbytes + 15) & ~15
    total += 16                                             # empty arrays still get a valid device address
    st = _STAGE.get(device)
    if st is None or st[0].numel() < total:
        st = [torch.empty(max(total, 1 << 16), dtype=torch.uint8, pin_memory=True), None]
        _STAGE[device] = st
    if st[1] is not None:
        st[1].synchronize()                                 # the previous copy out of the staging buffer has finished
    hv = st[0].numpy()
    for (_, a), o in zip(items, offs):
        if a.nbytes:
            hv[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
    dev = torch.empty(total, dtype=torch.uint8, device=device)
    dev.copy_(st[0][:total], non_blocking=True)
    st[1] = torch.cuda.Event()
    st[1].record(torch.cuda.current_stream(device))
    out = {}
    for (k, a), o in zip(items, offs):
        t = dev[o:o + a.nbytes].view(_TORCH_DTYPE[a.dtype.str])
        out[k] = t.view(a.shape) if a.ndim != 1 else t
    return out


_TORCH_DTYPE = {np.dtype(np.int32).str: torch.int32, np.dtype(np.int64).str: torch.int64, np.dtype(np.float32).str: torch.float32,
                np.dtype(np.float64).str: torch.float64, np.dtype(np.int16).str: torch.int16, np.dtype(np.uint8).str: torch.uint8,
                np.dtype(np.int8).str: torch.int8, np.dtype(np.bool_).str: torch.bool}


def upload_csr(csr, device):
    """Keeps the packed CSR of `pack_csr` resident on `device` (repeated extraction of the same raw batch, e.g. a
    benchmark loop, then skips the host -> device copies); returns `csr`."""
    csr["_device"] = _upload(csr, torch.device(device))
    return csr


def extract_batch(graphs, args, device="cuda"):
    """graphs: list of raw graphs (dicts as in kpgnn_b200.synth, or PyG-like Data objects);
    args = (K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count, max_distance_count, kernel).
    Returns a `kpgnn_b200.model.Batch` on `device` laid out as PyG's Batch.from_data_list would lay out the
    reference's per-graph results (node offsets applied, graph-major order)."""
    csr = pack_csr(graphs)
    xs = [g.get("x") if isinstance(g, dict) else _raw(g)[3] for g in graphs]
    extra = {"x": np.concatenate([np.asarray(v) for v in xs])} if xs and xs[0] is not None else None
    if extra is not None and extra["x"].dtype.str not in _TORCH_DTYPE:
        raise TypeError("unsupported node feature dtype %s" % extra["x"].dtype)
    csr["_device"] = dv = _upload(csr, device, extra)
    out = _extract_device(csr, *args, device=device)
    out.pop("eptr")
    return Batch(num_graphs=csr["G"], num_nodes=csr["N"], x=dv.get("x"), batch=dv["node_graph"].to(torch.int64), **out)


def extract_batch_host(graphs, args, device="cuda"):
    """extract_batch, then the fields copied to host memory (dict of CPU tensors + num_graphs/num_nodes)."""
    b = extract_batch(graphs, args, device)
    out = {f: (getattr(b, f).cpu() if torch.is_tensor(getattr(b, f)) else None) for f in Batch.FIELDS}
    out["num_graphs"], out["num_nodes"] = b.num_graphs, b.num_nodes
    return out


def extract_multi_hop_neighbors(data, K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count,
                                max_distance_count, kernel):
    """Reference-compatible per-graph entry point (data_utils.py:20-107): mutates and returns `data`."""
    edge_index, num_nodes = data.edge_index, data.num_nodes
    if edge_index.size(1) == 0:
        # graph with no edge: the reference returns early with these two fields only (note the second name and
        # its [N,K,max_hop_num] shape), data_utils.py:37-44
        data.peripheral_edge_attr = torch.zeros([num_nodes, K, max_edge_type, 2], dtype=torch.long)
        data.peripheral_configuration = torch.zeros([num_nodes, K, max_hop_num], dtype=torch.long)
        return data
    has_attr = ("edge_attr" in data) if hasattr(data, "__contains__") else getattr(data, "edge_attr", None) is not None
    g = {"num_nodes": num_nodes, "edge_index": edge_index.cpu().numpy(),
         "edge_attr": data.edge_attr.cpu().numpy() if has_attr else None}
    dev = edge_index.device if edge_index.is_cuda else torch.device("cuda")
    out = _extract_device(pack_csr([g]), K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count,
                          max_distance_count, kernel, dev)
    back = edge_index.device
    data.edge_index = out["edge_index"].to(back)
    data.edge_attr = out["edge_attr"].to(back)
    data.peripheral_edge_attr = None if out["peripheral_edge_attr"] is None else out["peripheral_edge_attr"].to(back)
    data.peripheral_configuration_attr = (None if out["peripheral_configuration_attr"] is None
                                          else out["peripheral_configuration_attr"].to(back))
    data.pe_attr = None if out["pe_attr"] is None else out["pe_attr"].to(back)
    return data


def extract_many(data_list, K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count, max_distance_count, kernel,
                 chunk=1024, device="cuda"):
    """extract_multi_hop_neighbors over a LIST of graphs -- what a dataset's process() does with its pre_transform
    (datasets/ZINC_dataset.py:100-140, PlanarSATPairsDataset.py:28-39: `[self.pre_transform(d) for d in data_list]`) -- with
    the graphs extracted `chunk` at a time in one batched GPU call instead of one call per graph.  Mutates and returns the
    same objects with exactly the fields, dtypes and quirks of the per-graph function (graphs without edges take the
    reference's early-return branch, data_utils.py:37-44)."""
    args = (K, max_edge_attr_num, max_hop_num, max_edge_type, max_edge_count, max_distance_count, kernel)
    todo = []
    for d in data_list:
        if d.edge_index.size(1) == 0:
            extract_multi_hop_neighbors(d, *args)                    # two zero fields, no kernel involved
        else:
            todo.append(d)
    for c0 in range(0, len(todo), chunk):
        part = todo[c0:c0 + chunk]
        raws = []
        for d in part:
            has_attr = ("edge_attr" in d) if hasattr(d, "__contains__") else getattr(d, "edge_attr", None) is not None
            raws.append({"num_nodes": d.num_nodes, "edge_index": d.edge_index.cpu().numpy(),
                         "edge_attr": d.edge_attr.cpu().numpy() if has_attr else None})
        csr = pack_csr(raws)
        csr["_device"] = _upload(csr, device)
        out = _extract_device(csr, *args, device=device)
        ei, ea = out["edge_index"].cpu(), out["edge_attr"].cpu()
        pea = None if out["peripheral_edge_attr"] is None else out["peripheral_edge_attr"].cpu()
        pca = None if out["peripheral_configuration_attr"] is None else out["peripheral_configuration_attr"].cpu()
        gptr = torch.from_numpy(csr["gptr"].astype(np.int64))
        eptr = torch.searchsorted(ei[0].contiguous(), gptr)         # K-hop edges are sorted by (graph, source, target)
        for g, d in enumerate(part):
            n0, n1, e0, e1 = int(gptr[g]), int(gptr[g + 1]), int(eptr[g]), int(eptr[g + 1])
            back = d.edge_index.device
            d.edge_index = (ei[:, e0:e1] - n0).to(back)
            d.edge_attr = ea[e0:e1].clone().to(back)
            d.peripheral_edge_attr = None if pea is None else pea[n0:n1].clone().to(back)
            d.peripheral_configuration_attr = None if pca is None else pca[n0:n1].clone().to(back)
            d.pe_attr = torch.zeros((n1 - n0, K - 1), dtype=torch.int64, device=back) if K > 1 else None
    return data_list


def post_transform(wo_path_encoding, wo_edge_feature):
    """Ablation clamps applied per access (data_utils.py:306-347): plain elementwise host ops, kept for import
    compatibility with the reference's train scripts."""
    def transform(g):
        ea = g.edge_attr
        if wo_edge_feature:
            ea[:, 0].clamp_(max=2)
        if wo_path_encoding:
            ea[:, 1:].clamp_(max=2)
        if wo_path_encoding and "pe_attr" in g:
            g.pe_attr.clamp_(max=0)
        g.edge_attr = ea
        return g
    if not (wo_path_encoding or wo_edge_feature):
        return lambda g: g
    return transform


def resistance_distance(data):
    """Optional `--use_rd` feature (data_utils.py:280-303): effective resistance to node 0 from the Laplacian
    pseudo-inverse.  Outside the K-hop hot path; dense host linear algebra as in the reference."""
    n = data.num_nodes
    ei = data.edge_index.cpu().numpy()
    adj = np.zeros((n, n))
    np.add.at(adj, (ei[0], ei[1]), 1.0)
    lap = np.diag(adj.sum(0)) - adj
    linv = np.linalg.pinv(lap)
    d = np.diag(linv)
    data.rd = torch.from_numpy((linv[0, 0] + d - linv[0, :] - linv[:, 0]).astype(np.float32)).unsqueeze(1)
    return data
