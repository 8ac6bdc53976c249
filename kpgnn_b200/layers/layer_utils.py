"""Layer factory -- mirror of the reference's layers/layer_utils.py:10-34 (same `args` fields, same result)."""
from .KPGCN import *  # noqa: F401,F403
from .KPGIN import *  # noqa: F401,F403
from .KPGINplus import *  # noqa: F401,F403
from .KPGraphSAGE import *  # noqa: F401,F403


def make_gnn_layer(args):
    name = args.model_name
    if name == "KPGCN":
        return KPGCNConv(args.hidden_size, args.hidden_size, args.K, args.num_hop1_edge, args.max_pe_num,
                         args.combine)
    if name in ("KPGIN", "KPGINPrime"):
        return KPGINConv(args.hidden_size, args.hidden_size, args.K, args.eps, args.train_eps,
                         args.num_hop1_edge, args.max_pe_num, args.combine)
    if name == "KPGraphSAGE":
        return KPGraphSAGEConv(args.hidden_size, args.hidden_size, args.K, args.aggr, args.num_hop1_edge,
                               args.max_pe_num, args.combine)
    if name == "KPGINPlus":
        # layer l (1-based) sees min(l, K) hops: the stack of previous layer outputs grows up to K
        return [KPGINPlusConv(args.hidden_size, args.hidden_size, min(l, args.K), args.num_hop1_edge,
                              args.max_pe_num, args.combine) for l in range(1, args.num_layer + 1)]
    raise ValueError("Not supported GNN type")
