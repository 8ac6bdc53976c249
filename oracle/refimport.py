"""TEST INFRASTRUCTURE ONLY: import the UNMODIFIED reference (read-only at /root/reference) behind the
`torch_geometric` stand-in in oracle/pyg_standin.  Only usable in the build container; on the GPU box the
reference does not exist and `available()` is False -- tests then fall back to the committed goldens."""
import importlib
import os
import sys

REF_ROOT = os.environ.get("KPGNN_REFERENCE_ROOT", "/root/reference")
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyg_standin")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "data_utils.py"))


def load():
    """Returns a namespace with the reference modules: data_utils, layers.*, models.*"""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    try:
        import torch_geometric  # noqa: F401  (a real install wins if one ever exists)
    except ImportError:
        if _STANDIN not in sys.path:
            sys.path.insert(0, _STANDIN)
    # the reference uses top-level names `layers`, `models`, `data_utils`; make sure they resolve to IT
    for name in list(sys.modules):
        if name in ("layers", "models", "data_utils") or name.startswith(("layers.", "models.")):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(REF_ROOT):
                del sys.modules[name]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class NS(object):
        pass
    ns = NS()
    ns.data_utils = importlib.import_module("data_utils")
    for m in ("KPGIN", "KPGINplus", "KPGCN", "KPGraphSAGE", "combine", "gine", "layer_utils",
              "feature_encoder", "input_encoder"):
        setattr(ns, m, importlib.import_module("layers." + m))
    ns.GNNs = importlib.import_module("models.GNNs")
    ns.GraphRegression = importlib.import_module("models.GraphRegression")
    ns.GraphClassification = importlib.import_module("models.GraphClassification")
    ns.model_utils = importlib.import_module("models.model_utils")
    import torch_geometric.data as pgd
    ns.Data, ns.Batch = pgd.Data, pgd.Batch
    return ns
