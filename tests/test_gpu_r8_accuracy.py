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


def _product_stack_run(tg, g, seed, dev, captured: bool):
    """One training run through the stack a user of the package runs: GCN.loss (fused forward + masked cross-entropy) or
    its CUDA-graph capture, tg.optim.Adam (tg_adam_f32), Philox dropout seeded from torch's generator, GCN.evaluate for the
    validation loss (early stopping) and the test metrics — the reference's loop, trainer.py:349-398."""
    from tests.golden.make_golden_shared import train_val_split
    n, nhid, nclass = int(g["n_docs"] + g["n_topics"]), int(g["nhid"]), int(g["nclass"])
    adj = torch.sparse_coo_tensor(torch.tensor(np.stack([g["adj_rows"], g["adj_cols"]]).astype(np.int64)),
                                  torch.tensor(g["adj_vals"]), (n, n), check_invariants=False).to(dev)
    x = tg.Featureless(n)
    target = torch.tensor(g["target"].astype(np.int64), device=dev)
    test_idx = torch.tensor(g["test"].astype(np.int64), device=dev)
    tr, va = train_val_split(g["train_all"], seed)
    tr_i, va_i = torch.tensor(tr, device=dev), torch.tensor(va, device=dev)
    torch.manual_seed(seed)
    model = tg.GCN(n, nhid, nclass, 0.5).to(dev)   # same initial weights as the reference for this seed; dropout seed drawn from torch
    opt = tg.optim.Adam(model.parameters(), lr=0.02)
    step = tg.CapturedTrainStep(model, x, adj, target, tr_i, warmup=1) if captured else None
    if captured:  # the capture's warm-up steps ran on the initial weights without an optimizer step: nothing to undo
        pass
    stop = EarlyStopping(10)
    epochs = 0
    for epoch in range(200):
        model.train()
        if captured:
            step.step()
        else:
            opt.zero_grad()
            model.loss(x, adj, target, tr_i).backward()
        opt.step()
        val = model.evaluate(x, adj, target, va_i, prefix="val")
        epochs += 1
        if stop(val["val_loss"]):
            break
    test = model.evaluate(x, adj, target, test_idx, prefix="test")
    return test["acc"], test["macro_f1"], epochs


def test_r8_accuracy_product_stack_20_seeds(r8_golden):
    """The DEFAULT production path (Philox dropout, fused loss, tg.optim.Adam, GCN.evaluate) trained end to end on the real
    R8 TopicGCN graph for 20 seeds, against 20 runs of the REAL reference loop with torch's own dropout stream
    (tests/golden/r8_training_unpaired.json, made by make_golden_r8_unpaired.py).  The masks differ, so the comparison is in
    distribution: mean test accuracy within 0.3 points (north star), every run inside the reference's range +- 1 point."""
    import topicgcn_b200 as tg
    ref = json.load(open(os.path.join(GOLDEN, "r8_training_unpaired.json")))
    ref_acc = np.array([r["test_acc"] for r in ref["runs"]])
    dev = torch.device("cuda:0")
    accs, f1s = [], []
    for seed in range(20):
        acc, f1, epochs = _product_stack_run(tg, r8_golden, seed, dev, captured=False)
        accs.append(acc); f1s.append(f1)
        assert ref_acc.min() - 0.01 <= acc <= ref_acc.max() + 0.01, (seed, acc, epochs)
    accs = np.array(accs)
    assert abs(accs.mean() - ref_acc.mean()) <= 0.003, (accs.mean(), ref_acc.mean(), accs.std(), ref_acc.std())
    ref_f1 = np.array([r["test_macro_f1"] for r in ref["runs"]])
    assert abs(np.mean(f1s) - ref_f1.mean()) <= 0.02, (np.mean(f1s), ref_f1.mean())


def test_r8_accuracy_captured_step_5_seeds(r8_golden):
    """The same loop with the train step replayed from a CUDA graph (CapturedTrainStep: device-side Philox counter)."""
    import topicgcn_b200 as tg
    ref = json.load(open(os.path.join(GOLDEN, "r8_training_unpaired.json")))
    ref_acc = np.array([r["test_acc"] for r in ref["runs"]])
    dev = torch.device("cuda:0")
    accs = [_product_stack_run(tg, r8_golden, seed, dev, captured=True)[0] for seed in range(5)]
    assert all(ref_acc.min() - 0.01 <= a <= ref_acc.max() + 0.01 for a in accs), accs
    assert abs(np.mean(accs) - ref_acc.mean()) <= 0.006, (accs, ref_acc.mean())   # 5 runs: twice the 20-run tolerance
