"""Stand-in for torch_geometric.utils (test infrastructure)."""
import numpy as np
import scipy.sparse
import torch


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    row, col = edge_index.cpu()
    if edge_attr is None:
        edge_attr = torch.ones(row.size(0))          # float32 ones, as PyG
    else:
        edge_attr = edge_attr.view(-1).cpu()
        assert edge_attr.size(0) == row.size(0)
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    return scipy.sparse.coo_matrix((edge_attr.numpy(), (row.numpy(), col.numpy())), (n, n))


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(0, n, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1), edge_attr


def from_networkx(G):
    from ..data import Data
    import networkx as nx
    G = nx.convert_node_labels_to_integers(G)
    G = G.to_directed() if not nx.is_directed(G) else G
    edges = list(G.edges)
    edge_index = torch.tensor(edges, dtype=torch.long).t().contiguous().view(2, -1)
    d = Data(edge_index=edge_index)
    d.num_nodes_ = G.number_of_nodes()
    return d
