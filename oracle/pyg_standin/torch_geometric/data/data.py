"""Stand-in for torch_geometric.data.{Data, Batch} (test infrastructure, see package docstring)."""
import torch


class Data(object):
    """Attribute bag.  None-valued attributes count as absent for `in` / `keys`, as in PyG."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.__dict__["_store"] = {}
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos).items():
            self._store[k] = v
        for k, v in kwargs.items():
            self._store[k] = v

    # pickles written by real PyG (data/EXP/raw/GRAPHSAT.pkl) carry plain instance __dict__ entries
    def __setstate__(self, state):
        self.__dict__["_store"] = {}
        src = state.get("_store", state) if isinstance(state, dict) else state
        if not isinstance(src, dict):
            src = getattr(src, "__dict__", {})
        for k, v in src.items():
            if k.startswith("_"):
                continue
            self._store[k] = v

    def __getstate__(self):
        return dict(self._store)

    def __getattr__(self, key):
        store = self.__dict__.get("_store")
        if store is None:
            self.__dict__["_store"] = store = {}
        if key in store:
            return store[key]
        if key.startswith("__"):
            raise AttributeError(key)
        raise AttributeError("Data has no attribute %r" % key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return self._store.get(key, None) is not None

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    @property
    def keys(self):
        return [k for k, v in self._store.items() if v is not None]

    @property
    def num_nodes(self):
        if self._store.get("num_nodes_", None) is not None:
            return self._store["num_nodes_"]
        x = self._store.get("x", None)
        if x is not None:
            return x.size(0)
        ei = self._store.get("edge_index", None)
        if ei is not None and ei.numel() > 0:
            return int(ei.max()) + 1
        return 0

    @property
    def num_edges(self):
        ei = self._store.get("edge_index", None)
        return 0 if ei is None else ei.size(1)

    def to(self, device):
        for k, v in self._store.items():
            if torch.is_tensor(v):
                self._store[k] = v.to(device)
        return self

    def clone(self):
        out = self.__class__()
        for k, v in self._store.items():
            out._store[k] = v.clone() if torch.is_tensor(v) else v
        return out


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        """Concatenate every field along dim 0; `edge_index` along dim -1 with cumulative node offsets."""
        out = cls()
        keys = []
        for d in data_list:
            for k in d.keys:
                if k not in keys and k != "num_nodes_":
                    keys.append(k)
        offset = 0
        cols = {k: [] for k in keys}
        batch = []
        for i, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d._store.get(k, None)
                if v is None:
                    continue
                if k == "edge_index":
                    v = v + offset
                elif not torch.is_tensor(v):
                    v = torch.tensor([v])
                elif v.dim() == 0:
                    v = v.view(1)
                cols[k].append(v)
            batch.append(torch.full((n,), i, dtype=torch.long))
            offset += n
        for k in keys:
            if not cols[k]:
                continue
            out._store[k] = torch.cat(cols[k], dim=-1 if k == "edge_index" else 0)
        out._store["batch"] = torch.cat(batch) if batch else torch.zeros(0, dtype=torch.long)
        out._store["num_graphs"] = len(data_list)
        out._store["num_nodes_"] = offset
        return out
