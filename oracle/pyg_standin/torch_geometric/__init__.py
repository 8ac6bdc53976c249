"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the un-vendored dependency `torch_geometric`.

The reference (JiaruiFeng/KP-GNN, README.md:9-14) pins PyG=2.1.0, which is not installed in this image and
has no wheel in the offline wheelhouse.  This package provides exactly the API surface the reference's
hot path touches so that `/root/reference/{data_utils,layers,models}` can be imported
UNMODIFIED in the build container to (a) validate the oracle restatements under `oracle/` and (b) generate
the golden vectors under `tests/golden/`.  It implements the documented PyG semantics: source->target flow,
`x_j = x[edge_index[0]]`, reduction at `edge_index[1]`, `dim_size = N`.

Nothing in the product package (`kpgnn_b200/`) imports this.
"""
__version__ = "2.1.0-standin"
from . import data, utils, nn, loader  # noqa: F401
