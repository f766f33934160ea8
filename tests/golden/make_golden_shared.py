"""Pieces shared by make_golden.py (build container, real reference) and the tests (any box, no reference):
how seeds map to initial parameters, splits and dropout masks in the paired runs."""
import math

import numpy as np


def mask_seed(seed: int, epoch: int) -> int:
    """Seed of torch's CPU generator right before the train-mode forward of `epoch`."""
    return 1_000_003 * (seed + 1) + epoch


def reference_init(seed: int, nfeat: int, nhid: int, nclass: int) -> dict:
    """Initial parameters of `layer.GCN(nfeat, nhid, nclass, p)` after th.manual_seed(seed): the reference draws
    gc1.weight, gc1.bias, gc2.weight, gc2.bias in that order, each U(-1/sqrt(out), 1/sqrt(out))
    (reference layer.py:67-82, 155-159)."""
    import torch
    torch.manual_seed(seed)
    out = {}
    for name, (i, o) in (("gc1", (nfeat, nhid)), ("gc2", (nhid, nclass))):
        stdv = 1.0 / math.sqrt(o)
        out[f"{name}.weight"] = torch.empty(i, o).uniform_(-stdv, stdv).numpy()
        out[f"{name}.bias"] = torch.empty(o).uniform_(-stdv, stdv).numpy()
    return out


def train_val_split(train_all, seed: int, val_ratio: float = 0.1):
    """trainer.py:335-338: sklearn train_test_split(train_lst, test_size=val_ratio, shuffle=True, random_state=seed)."""
    from sklearn.model_selection import train_test_split
    tr, va = train_test_split(np.asarray(train_all).tolist(), test_size=val_ratio, shuffle=True, random_state=seed)
    return np.asarray(tr, dtype=np.int64), np.asarray(va, dtype=np.int64)
