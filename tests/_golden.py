"""Helpers to load the fixtures written by oracle/gen_golden.py."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    """-> dict(meta=..., sd={..}, inp={..}, out={..}, other={prefix: {..}}) with torch tensors."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    res = {"meta": json.loads(str(z["meta"])), "sd": {}, "inp": {}, "out": {}, "other": {}}
    for key in z.files:
        if key == "meta":
            continue
        prefix, rest = key.split("/", 1)
        t = torch.from_numpy(np.array(z[key]))
        if prefix == "sd":
            res["sd"][rest] = t
        elif prefix == "in":
            res["inp"][rest] = t
        elif prefix == "out":
            res["out"][rest] = t
        else:
            res["other"].setdefault(prefix, {})[rest] = t
    if "layers" in res["meta"]:
        res["meta"]["layers"] = [tuple(l) for l in res["meta"]["layers"]]
    return res


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rel_linf(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
