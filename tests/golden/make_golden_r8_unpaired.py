#!/usr/bin/env python
"""Unpaired reference distribution for the product-stack accuracy test (tests/test_gpu_r8_accuracy.py).

Runs the REAL reference modules (`/root/reference/layer.py`, `utils.py`: build container only) on the real R8 TopicGCN
graph stored in r8_topic.npz with the reference's training loop (trainer.py:298-398: Adam lr 0.02, dropout 0.5, early
stopping patience 10 on the validation loss, <= 200 epochs) for 20 seeds.  Unlike r8_training.json the dropout masks are
NOT re-seeded per epoch: every run uses torch's global generator as the reference does (th.manual_seed(seed) once,
trainer.py:294-296), because the product stack draws its masks from its own Philox stream and can only be compared
with the reference in distribution.  Output: r8_training_unpaired.json (seed, epochs, test accuracy, macro F1)."""
import json
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
np.Inf = np.inf  # numpy 2 removed the alias utils.py:234 uses
sys.modules.setdefault("prettytable", types.SimpleNamespace(PrettyTable=object))  # utils.py:171 imports it for printing only
import layer as ref_layer  # noqa: E402
import utils as ref_utils  # noqa: E402
from tests.golden.make_golden_shared import train_val_split  # noqa: E402


def main(n_seeds: int = 20):
    g = np.load(os.path.join(HERE, "r8_topic.npz"))
    n = int(g["n_docs"] + g["n_topics"])
    adj = th.sparse_coo_tensor(th.tensor(np.stack([g["adj_rows"], g["adj_cols"]]).astype(np.int64)), th.tensor(g["adj_vals"]), (n, n))
    ar = th.arange(n)
    x = th.sparse_coo_tensor(th.stack([ar, ar]), th.ones(n), (n, n))
    target = th.tensor(g["target"].astype(np.int64))
    te_i = th.tensor(g["test"].astype(np.int64))
    crit = th.nn.CrossEntropyLoss()
    runs = []
    for seed in range(n_seeds):
        tr, va = train_val_split(g["train_all"], seed)
        tr_i, va_i = th.tensor(tr), th.tensor(va)
        th.manual_seed(seed)
        np.random.seed(seed)
        model = ref_layer.GCN(nfeat=n, nhid=int(g["nhid"]), nclass=int(g["nclass"]), dropout=0.5)
        opt = th.optim.Adam(model.parameters(), lr=0.02)
        stopper = ref_utils.EarlyStopping(10)
        epochs = 0
        for epoch in range(200):
            model.train()
            opt.zero_grad()
            logits = model.forward(x, adj)
            loss = crit(logits[tr_i], target[tr_i])
            loss.backward()
            opt.step()
            model.eval()
            with th.no_grad():
                vloss = float(crit(model.forward(x, adj)[va_i], target[va_i]).item())
            epochs += 1
            if stopper(vloss):
                break
        model.eval()
        with th.no_grad():
            lg = model.forward(x, adj)
            acc = ref_utils.accuracy(lg[te_i], target[te_i])
            f1 = ref_utils.macro_f1(lg[te_i], target[te_i], int(g["nclass"]))
        f1v = float(f1[0]) if isinstance(f1, (tuple, list)) else float(f1)
        runs.append({"seed": seed, "epochs": epochs, "test_acc": float(acc), "test_macro_f1": f1v})
        print(f"seed {seed}: {epochs} epochs, test acc {acc:.4f}", flush=True)
    with open(os.path.join(HERE, "r8_training_unpaired.json"), "w") as fh:
        json.dump({"config": {"nhid": int(g["nhid"]), "dropout": 0.5, "lr": 0.02, "max_epoch": 200, "early_stopping": 10,
                              "val_ratio": 0.1, "featureless": True, "torch": th.__version__,
                              "dropout_stream": "torch global generator, seeded once per run (unpaired)"},
                   "runs": runs}, fh, indent=0)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
