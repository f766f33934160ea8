"""Config 1 (BASELINE.json): the real R8 TopicGCN graph, featureless, trained with the reference's loop
(trainer.py:298-398: Adam lr 0.02, dropout 0.5, early stopping patience 10 on validation loss, <= 200 epochs) for the 5
pinned seeds of tests/golden/r8_training.json, on the SAME initial weights, splits and dropout masks the real reference
used.  North-star criterion: test accuracy within 0.3 points over 5 seeds."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class EarlyStopping:
    """utils.EarlyStopping (reference utils.py:216-255): patience on validation loss."""

    def __init__(self, patience):
        self.patience, self.counter, self.best = patience, 0, None

    def __call__(self, val_loss):
        score = -val_loss
        if self.best is None:
            self.best = score
        elif score < self.best:
            self.counter += 1
            if self.counter >= self.patience:
                return True
        else:
            self.best, self.counter = score, 0
        return False


def test_r8_accuracy_parity_5_seeds(r8_golden):
    import topicgcn_b200 as tg
    from tests.golden.make_golden_shared import mask_seed, train_val_split
    g = r8_golden
    ref = json.load(open(os.path.join(GOLDEN, "r8_training.json")))
    dev = torch.device("cuda:0")
    n, nd, nhid, nclass = int(g["n_docs"] + g["n_topics"]), int(g["n_docs"]), int(g["nhid"]), int(g["nclass"])
    adj = torch.sparse_coo_tensor(torch.tensor(np.stack([g["adj_rows"], g["adj_cols"]]).astype(np.int64)),
                                  torch.tensor(g["adj_vals"]), (n, n), check_invariants=False).to(dev)
    x = tg.Featureless(n)
    target = torch.tensor(g["target"].astype(np.int64), device=dev)
    test_idx = torch.tensor(g["test"].astype(np.int64), device=dev)
    crit = torch.nn.CrossEntropyLoss()
    accs, ref_accs = [], []
    for run in ref["runs"]:
        seed = run["seed"]
        tr, va = train_val_split(g["train_all"], seed)
        tr_i, va_i = torch.tensor(tr, device=dev), torch.tensor(va, device=dev)
        torch.manual_seed(seed)
        model = tg.GCN(n, nhid, nclass, 0.5).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=0.02)
        stop = EarlyStopping(10)
        epochs = 0
        for epoch in range(200):
            model.train()
            opt.zero_grad()
            torch.manual_seed(mask_seed(seed, epoch))  # the mask th.dropout drew in the reference run
            model.set_next_dropout_mask(torch.empty(n, nhid).bernoulli_(0.5).to(torch.uint8).to(dev))
            logits = model.forward(x, adj)
            loss = crit(logits[tr_i], target[tr_i])  # the reference call site, unchanged
            loss.backward()
            opt.step()
            model.eval()
            with torch.no_grad():
                lg = model.forward(x, adj)
                vloss = float(crit(lg[va_i], target[va_i]))
            epochs += 1
            if epoch < 5:  # the first epochs follow the reference trajectory closely
                assert abs(float(loss) - run["history"][epoch]["train_loss"]) <= 2e-3, (seed, epoch)
            if stop(vloss):
                break
        model.eval()
        with torch.no_grad():
            lg = model.forward(x, adj)
            acc = float((lg[test_idx].argmax(1) == target[test_idx]).float().mean())
        accs.append(acc)
        ref_accs.append(run["test_acc"])
        assert abs(acc - run["test_acc"]) <= 0.015, (seed, acc, run["test_acc"], epochs, run["epochs"])
    assert abs(np.mean(accs) - np.mean(ref_accs)) <= 0.003, (accs, ref_accs)  # within 0.3 points over 5 seeds
