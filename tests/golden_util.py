"""Readers for the committed reference outputs under tests/golden/ (written by oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else None
    return z, meta


def layer_case(z, idx):
    pre = "l%d_" % idx
    out = {"sd": {}, "gp": {}}
    for k in z.files:
        if not k.startswith(pre):
            continue
        name = k[len(pre):]
        v = torch.from_numpy(z[k])
        if name.startswith("sd_"):
            out["sd"][name[3:]] = v
        elif name.startswith("gp_"):
            out["gp"][name[3:]] = v
        else:
            out[name] = v
    return out


def build_layer(ctor, namespace):
    """ctor = [class name, *args]; namespace maps class names to constructors (oracle or product)."""
    return namespace[ctor[0]](*ctor[1:])
