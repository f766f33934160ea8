"""Golden vectors for the validation metrics, produced by the REAL reference (utils.accuracy / utils.macro_f1,
/root/reference/utils.py:25-109) in the build container:  python tests/golden/make_golden_metrics.py
Writes tests/golden/metrics.npz: logits, targets and the reference's (acc, f1, precision, recall) for a few cases,
including classes that never occur in the targets / predictions (the 0/0 -> 0 rule of utils.py:73-79)."""
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
if not hasattr(np, "Inf"):
    np.Inf = np.inf  # utils.py:234 (numpy 2 removed the alias)
if "prettytable" not in sys.modules:  # utils.py:171 imports it; not installed here and not needed for the metrics
    stub = types.ModuleType("prettytable")
    stub.PrettyTable = object
    sys.modules["prettytable"] = stub
import utils  # noqa: E402  (the reference module)


def main():
    rng = np.random.default_rng(0)
    out = {}
    cases = [(500, 8, 8), (2000, 20, 20), (64, 5, 3), (300, 4, 4)]  # (rows, classes, classes that occur in the targets)
    for i, (n, c, c_used) in enumerate(cases):
        logits = rng.normal(size=(n, c)).astype(np.float32)
        targ = rng.integers(0, c_used, size=n)
        logits[np.arange(n), targ] += rng.choice([0.0, 2.5], size=n).astype(np.float32)  # ~60 % correct
        if i == 3:
            logits[:, 3] = -50.0  # class 3 is never predicted: precision 0/0
        acc = utils.accuracy(th.tensor(logits), th.tensor(targ))
        f1, prec, rec = utils.macro_f1(th.tensor(logits), th.tensor(targ), num_classes=c)
        out[f"logits_{i}"] = logits
        out[f"target_{i}"] = targ.astype(np.int64)
        out[f"ref_{i}"] = np.array([acc, f1, prec, rec], dtype=np.float64)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("wrote metrics.npz", {k: v.shape for k, v in out.items() if k.startswith("ref")})


if __name__ == "__main__":
    main()
