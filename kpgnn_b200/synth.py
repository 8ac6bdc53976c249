"""Synthetic workload generators: ZINC-shaped molecules, typed random graphs, r-regular
graphs.  Pure numpy, deterministic in `seed`; used by bench.py, the tests and the golden-vector script.

A "raw graph" is a dict  {num_nodes, x [n] int64, edge_index [2,E] int64, edge_attr [E] int64 or None}
mirroring the fields of the reference's PyG `Data` before `extract_multi_hop_neighbors` (data_utils.py:20).
"""
import numpy as np


def _tree_dist(adj, s, cutoff):
    dist = {s: 0}
    frontier = [s]
    for d in range(1, cutoff + 1):
        nxt = []
        for u in frontier:
            for v in adj[u]:
                if v not in dist:
                    dist[v] = d
                    nxt.append(v)
        frontier = nxt
    return dist


def zinc_like_graph(rng):
    """One ZINC-shaped molecule: ~23 atoms, ~25 bonds (chain-biased tree of max degree 3 plus a few 5/6-ring
    closures, max degree 4), bond types {1,2,3} stored +1 as in train_ZINC.py:96-99, atom types in [0,21)."""
    n = int(np.clip(np.rint(rng.normal(23.2, 4.5)), 9, 37))
    adj = [[] for _ in range(n)]
    for v in range(1, n):
        # chain bias: attach to the previous atom most of the time, else to a random earlier atom
        for _ in range(16):
            u = v - 1 if rng.random() < 0.7 else int(rng.integers(0, v))
            if len(adj[u]) < 3:
                break
        else:
            u = min(range(v), key=lambda t: len(adj[t]))
        adj[u].append(v)
        adj[v].append(u)
    n_rings = int(max(0, np.rint(rng.normal(2.7, 1.0))))
    for _ in range(n_rings):
        for _try in range(20):
            a = int(rng.integers(0, n))
            if len(adj[a]) >= 4:
                continue
            dist = _tree_dist(adj, a, 5)
            cand = [v for v, d in dist.items() if d in (4, 5) and len(adj[v]) < 4 and v not in adj[a]]
            if cand:
                b = cand[int(rng.integers(0, len(cand)))]
                adj[a].append(b)
                adj[b].append(a)
                break
    src, dst, typ = [], [], []
    for u in range(n):
        for v in adj[u]:
            if u < v:
                t = int(rng.choice([1, 2, 3], p=[0.72, 0.25, 0.03])) + 1
                src += [u, v]
                dst += [v, u]
                typ += [t, t]
    order = np.lexsort((np.array(dst), np.array(src)))
    return {
        "num_nodes": n,
        "x": rng.integers(0, 21, size=n).astype(np.int64),
        "edge_index": np.stack([np.array(src)[order], np.array(dst)[order]]).astype(np.int64),
        "edge_attr": np.array(typ)[order].astype(np.int64),
        "y": float(rng.normal()),
    }


def zinc_like_graphs(num_graphs, seed=0):
    rng = np.random.default_rng(seed)
    return [zinc_like_graph(rng) for _ in range(num_graphs)]


def random_typed_graph(rng, n, p, num_types=3, directed=False, typed=True):
    """G(n,p) with edge types in [2, 2+num_types); `directed` draws each direction independently."""
    m = rng.random((n, n)) < p
    np.fill_diagonal(m, False)
    if not directed:
        m = np.triu(m, 1)
        m = m | m.T
    t = rng.integers(2, 2 + num_types, size=(n, n))
    if not directed:
        t = np.triu(t, 1)
        t = t + t.T
    src, dst = np.nonzero(m)
    return {
        "num_nodes": n,
        "x": np.zeros(n, dtype=np.int64),
        "edge_index": np.stack([src, dst]).astype(np.int64),
        "edge_attr": t[src, dst].astype(np.int64) if typed else None,
        "y": 0.0,
    }


def regular_graph(n, r=3, seed=0):
    """Random r-regular graph as run_simulation.py:119-129 builds it (networkx generator, x = ones)."""
    import networkx as nx
    g = nx.random_regular_graph(d=r, n=n, seed=seed)
    e = np.array(list(g.edges), dtype=np.int64).reshape(-1, 2)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    order = np.lexsort((dst, src))
    return {
        "num_nodes": n,
        "x": np.ones(n, dtype=np.int64),
        "edge_index": np.stack([src[order], dst[order]]).astype(np.int64),
        "edge_attr": None,
        "y": 0.0,
    }
