"""Stand-in for torch_geometric.data (test infrastructure, see package docstring)."""
from .data import Data, Batch  # noqa: F401
