"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's K-hop / peripheral
extraction, `/root/reference/data_utils.py:20-241`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.
It is the CHECKER for the CUDA extractor in `kpgnn_b200/csrc/extract.cu`; nothing in `kpgnn_b200/` imports it.

Parity status: PINNED against the reference itself, executed unmodified in the build container behind the
`torch_geometric` stand-in (oracle/refimport.py): see tests/test_oracle_cpu.py (live comparison when
/root/reference exists) and tests/golden/extract.npz (committed outputs of the reference, made by
oracle/make_golden.py).  The reference has no tests or golden vectors of its own.

Arithmetic domain: the reference computes walk counts with float32 sparse matmuls and casts to int32
(data_utils.py:117-120).  This restatement uses int64 walk counts saturated at SAT = 2**31 - 1.  Outputs agree
whenever every true walk count (and, for gd, the per-pair sum over hops) is < 2**31 and every cap is < 2**24;
beyond that the reference's float->int32 cast is undefined behaviour and there is nothing to be exact against.
"""
import numpy as np

SAT = 2 ** 31 - 1


def dense_inputs(num_nodes, edge_index, edge_attr=None):
    """A[u,v] = multiplicity of edge u->v, T[u,v] = summed edge-type value (data_utils.py:46-53: the COO ->
    dense conversions sum duplicate entries; missing edge_attr defaults to type 2)."""
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    if edge_attr is None:
        et = np.full(src.shape, 2, dtype=np.int64)
    else:
        et = np.asarray(edge_attr, dtype=np.int64).reshape(-1)
    A = np.zeros((num_nodes, num_nodes), dtype=np.int64)
    T = np.zeros((num_nodes, num_nodes), dtype=np.int64)
    np.add.at(A, (src, dst), 1)
    np.add.at(T, (src, dst), et)
    return A, T


def walk_counts(A, K):
    """W[k] = A^(k+1) with saturation, diagonal cleared AFTER all powers are taken (data_utils.py:110-125)."""
    out = []
    cur = A.copy()
    for k in range(K):
        if k > 0:
            cur = np.minimum(cur @ A, SAT)
        out.append(cur)
    res = []
    for w in out:
        w = w.copy()
        np.fill_diagonal(w, 0)
        res.append(w)
    return res


def hop_matrices(A, K, kernel):
    """Per-hop count matrices as used downstream.  gd: raw walk counts (data_utils.py:57-62);
    spd: counts masked to pairs first reached at that hop (data_utils.py:63-74)."""
    W = walk_counts(A, K)
    if kernel == "gd":
        member = np.zeros_like(A, dtype=bool)
        for w in W:
            member |= w > 0
        return W, member
    seen = W[0] > 0
    hops = [W[0]]
    for k in range(1, K):
        w = np.where(seen, 0, W[k])
        seen = seen | (w > 0)
        hops.append(w)
    return hops, seen


def _induced_distances(sub_adj, max_hop):
    """All-pairs directed BFS distance inside an induced subgraph, cutoff max_hop, 0 for self / unreachable
    (data_utils.py:224-241)."""
    m = sub_adj.shape[0]
    D = np.zeros((m, m), dtype=np.int64)
    reach = np.eye(m, dtype=bool)
    front = reach.copy()
    a = sub_adj.astype(np.int64)
    for h in range(1, max_hop + 1):
        nxt = ((front.astype(np.int64) @ a) > 0) & ~reach
        if not nxt.any():
            break
        D[nxt] = h
        reach |= nxt
        front = nxt
    return D


def peripheral_one_hop(T, Wk, max_hop_num, max_edge_type, max_edge_count, max_distance_count):
    """data_utils.py:165-221 for one hop.  Returns ([N, max_edge_type, 2], [N, max_hop_num+1]) int64."""
    n = T.shape[0]
    pe = np.zeros((n, max_edge_type, 2), dtype=np.int64)
    pc = np.zeros((n, max_hop_num + 1), dtype=np.int64)
    for i in range(n):
        S = np.nonzero(Wk[i] > 0)[0]                       # :185
        if S.size < 2:                                     # :188
            continue
        sub = T[np.ix_(S, S)]                              # :190
        w = sub[sub != 0]                                  # directed edges of the induced subgraph, :191-192
        if w.size == 0:                                    # :193
            continue
        cnt = np.bincount(w, minlength=max_edge_type + 2)[2:]            # :196-198
        order = np.argsort(-cnt, kind="stable")[:max_edge_type]         # :199-201 (ties: ascending index)
        pe[i, :, 0] = order
        pe[i, :, 1] = np.minimum(cnt[order], max_edge_count)            # :202
        D = _induced_distances(sub != 0, max_hop_num)                    # :205
        total = 0
        for h in range(1, max_hop_num + 1):                              # :207-214
            M = (D == h)
            big = M.sum(1) >= 2
            if big.any():
                Mi = M.astype(np.int64)
                total += int((((Mi @ sub) * Mi).sum(1) * big).sum())
        cf = np.bincount(D.reshape(-1), minlength=max_hop_num + 1)       # :216
        cf[0] = total                                                    # :218
        pc[i] = np.minimum(cf, max_distance_count)                       # :219
    return pe, pc


def extract_multi_hop_neighbors_np(num_nodes, edge_index, edge_attr, K, max_edge_attr_num, max_hop_num,
                                   max_edge_type, max_edge_count, max_distance_count, kernel):
    """data_utils.py:20-107 on plain arrays.  Returns a dict whose keys are the `Data` fields the reference
    sets (None-valued fields are returned as None; fields it leaves untouched are absent)."""
    edge_index = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    if edge_index.shape[1] == 0:                                         # :37-44 (note the field names/shapes)
        return {
            "peripheral_edge_attr": np.zeros((num_nodes, K, max_edge_type, 2), dtype=np.int64),
            "peripheral_configuration": np.zeros((num_nodes, K, max_hop_num), dtype=np.int64),
        }
    A, T = dense_inputs(num_nodes, edge_index, edge_attr)
    hops, member = hop_matrices(A, K, kernel)
    src, dst = np.nonzero(member)                                        # :76-78, row-major
    cols = [T[src, dst]]                                                 # :80-81
    for k in range(1, K):                                                # :83-90
        a = np.minimum(hops[k], max_edge_attr_num)
        a = np.where(a > 0, a + 1, a)
        cols.append(a[src, dst])
    out = {
        "edge_index": np.stack([src, dst]).astype(np.int64),
        "edge_attr": np.stack(cols, axis=1).astype(np.int64),
        "pe_attr": np.zeros((num_nodes, K - 1), dtype=np.int64) if K > 1 else None,   # :91-96, diag is 0
    }
    if max_hop_num > 0 and max_edge_type > 0:                            # :141
        pes, pcs = [], []
        for k in range(K):
            pe, pc = peripheral_one_hop(T, hops[k], max_hop_num, max_edge_type, max_edge_count,
                                        max_distance_count)
            pes.append(pe)
            pcs.append(pc)
        out["peripheral_edge_attr"] = np.stack(pes, axis=1)              # [N, K, max_edge_type, 2], :154-157
        out["peripheral_configuration_attr"] = np.stack(pcs, axis=1)     # [N, K, max_hop_num+1]
    else:
        out["peripheral_edge_attr"] = None
        out["peripheral_configuration_attr"] = None
    return out
