"""TEST INFRASTRUCTURE ONLY: import the UNMODIFIED reference behind the `torch_geometric` stand-in in
oracle/pyg_standin.  The reference is read from /root/reference in the build container and from the git-ignored
staging copy oracle/_ref/ (made by oracle/fetch_ref.py, travels with gpurun) on the GPU box; when neither exists
`available()` is False and tests fall back to the committed goldens."""
import importlib
import os
import sys

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("KPGNN_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/data_utils.py") else _STAGED)
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


def load_models_over_dropin():
    """The reference's models/GNNs.py, UNMODIFIED, imported a second time with its `from layers... import` lines
    resolving to the product's drop-in package (what kpgnn_b200.install_dropin() does for a user's process), under
    private module names so that it coexists with the all-reference namespace of load().  Returns a namespace with
    GNNs, GraphRegression, GraphClassification built on the sm_100a layers."""
    import importlib.util
    import kpgnn_b200
    ns_ref = load()                                    # stand-in torch_geometric on sys.path, REF_ROOT importable
    saved = {k: v for k, v in sys.modules.items()
             if k in ("layers", "models", "data_utils") or k.startswith(("layers.", "models."))}
    for k in saved:
        del sys.modules[k]
    try:
        kpgnn_b200.install_dropin()

        class NS(object):
            pass
        out = NS()
        for name in ("GNNs", "GraphRegression", "GraphClassification"):
            spec = importlib.util.spec_from_file_location("kp_dropin_models_" + name,
                                                          os.path.join(REF_ROOT, "models", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            setattr(out, name, mod)
        assert out.GNNs.GINEConv.__module__.startswith("kpgnn_b200."), out.GNNs.GINEConv.__module__
        out.ref = ns_ref
        return out
    finally:
        for k in list(sys.modules):
            if k in ("layers", "models", "data_utils") or k.startswith(("layers.", "models.")):
                del sys.modules[k]
        sys.modules.update(saved)
