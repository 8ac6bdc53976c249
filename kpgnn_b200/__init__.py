"""kpgnn_b200 -- B200 (sm_100a) implementation of KP-GNN's K-hop extraction + aggregation path.

    kpgnn_b200.layers.*      drop-in mirrors of the reference's layers/ modules
    kpgnn_b200.data_utils    drop-in mirror of the reference's data_utils.py (GPU extractor)
    kpgnn_b200.ops / plan    the operator and the graph plan the layers are built on
    kpgnn_b200.model         plain-torch caller (KP-GIN+ regressor) used by bench.py / smoke()
    kpgnn_b200.build         nvcc build of the C-ABI library (include/kpgnn.h)
"""
import importlib
import sys

__version__ = "0.1.0"

_DROPIN_LAYER_MODULES = ("KPGCN", "KPGIN", "KPGINplus", "KPGraphSAGE", "combine", "gine", "layer_utils",
                         "feature_encoder", "input_encoder")


def install_dropin():
    """Make the reference's import names resolve to this package: after the call, `from layers.gine import
    GINEConv`, `from layers.layer_utils import make_gnn_layer`, `from data_utils import
    extract_multi_hop_neighbors` (models/GNNs.py:9-10, train_ZINC.py:19) import the sm_100a implementations,
    so the reference's `models/` and train scripts run on them unchanged."""
    pkg = importlib.import_module("kpgnn_b200.layers")
    sys.modules["layers"] = pkg
    for name in _DROPIN_LAYER_MODULES:
        sys.modules["layers." + name] = importlib.import_module("kpgnn_b200.layers." + name)
    sys.modules["data_utils"] = importlib.import_module("kpgnn_b200.data_utils")
    return pkg
