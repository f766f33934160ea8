"""Generate the golden fixtures under tests/golden/ by running the REAL reference code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [--r8-graph /tmp/ref_work/data/graph/R8_topic.txt]

It imports the unmodified reference modules `layer` (GCN / GraphConvolution) and `utils` (preprocess_adj) from
/root/reference and records their inputs and outputs:

  small_dtt.npz    a small synthetic document-topic-topic graph pushed through the reference's own ingest
                   (networkx graph -> nx.adjacency_matrix -> symmetrise trainer.py:148 -> utils.preprocess_adj),
                   reference GCN parameters for a fixed seed, eval/train logits, the dropout keep mask, loss and all
                   four parameter gradients from loss.backward(); featureless (X = I) and sparse-feature variants.
  r8_topic.npz     the real R8 TopicGCN graph (edge list written by the reference's build_graph.py, 50 topics,
                   --no_word2vec), its normalised adjacency from the reference ingest, labels/splits from
                   data/text_dataset/R8.txt, reference logits/loss/gradient checks for seed 0.
  r8_training.json the reference model trained featureless on that graph for 5 pinned seeds with the reference's
                   loop (trainer.py:349-376: Adam lr 0.02, dropout 0.5, early stopping 10 on val loss); dropout masks
                   are re-seeded per epoch so that another implementation can be run on the SAME masks.

Shims needed to import the reference on the container's stack: numpy 2 removed np.Inf.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

if not hasattr(np, "Inf"):
    np.Inf = np.inf  # reference utils.py:234

REF = "/root/reference"
sys.path.insert(0, REF)

import networkx as nx  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch as th  # noqa: E402

import layer as ref_layer  # noqa: E402  (reference layer.py)
import utils as ref_utils  # noqa: E402  (reference utils.py)

HERE = os.path.dirname(os.path.abspath(__file__))


sys.path.insert(0, HERE)
from make_golden_shared import mask_seed  # noqa: E402  (seed -> dropout-mask seed, shared with the tests)


def reference_ingest(graph: nx.Graph):
    """trainer.py:98-151: adjacency from networkx, symmetrise, preprocess_adj -> torch sparse COO."""
    n = graph.number_of_nodes()
    adj = nx.adjacency_matrix(graph, nodelist=list(range(n)), weight="weight", dtype=np.float32)
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)
    raw = sp.coo_matrix(adj)
    return ref_utils.preprocess_adj(adj, is_sparse=True), raw


def sparse_identity(n: int):
    idx = th.arange(n, dtype=th.int64)
    return th.sparse_coo_tensor(th.stack([idx, idx]), th.ones(n), (n, n))


def run_reference(model, x, adj, target, index, train: bool, mseed: int):
    """One forward (+ loss + backward in train mode) of the reference model; returns numpy results."""
    model.train(train)
    model.zero_grad()
    th.manual_seed(mseed)
    logits = model.forward(x, adj)
    loss = th.nn.CrossEntropyLoss()(logits[index], target[index])
    out = {"logits": logits.detach().numpy().copy(), "loss": float(loss.item())}
    if train:
        loss.backward()
        out["grads"] = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}
    return out


def drawn_mask(mseed: int, n: int, h: int, p: float) -> np.ndarray:
    """The Bernoulli(1-p) sample th.dropout draws for an [n x h] input after th.manual_seed(mseed)."""
    th.manual_seed(mseed)
    return th.empty(n, h).bernoulli_(1 - p).numpy().astype(np.uint8)


def small_graph(seed: int = 0, n_docs: int = 300, n_topics: int = 12):
    rng = np.random.default_rng(seed)
    g = nx.Graph()
    g.add_nodes_from(range(n_docs + n_topics))
    for d in range(n_docs):
        deg = int(rng.integers(2, 7))
        topics = rng.choice(n_topics, size=deg, replace=False, p=zipf_p(n_topics))
        w = rng.random(deg)
        w = w / w.sum()
        for t, wt in zip(topics, w):
            if wt >= 0.02:  # build_graph.py:105-107
                g.add_edge(d, n_docs + int(t), weight=float(wt))
    emb = rng.normal(size=(n_topics, 16)) + 0.8
    sim = emb @ emb.T / np.outer(np.linalg.norm(emb, axis=1), np.linalg.norm(emb, axis=1))
    for i in range(n_topics):
        for j in range(i + 1, n_topics):
            if sim[i, j] > 0.3:  # build_graph.py:125
                g.add_edge(n_docs + i, n_docs + j, weight=float(sim[i, j]))
    return g, n_docs, n_topics


def zipf_p(k: int):
    p = 1.0 / np.arange(1, k + 1)
    return p / p.sum()


def make_small():
    g, n_docs, n_topics = small_graph()
    n = n_docs + n_topics
    adj, raw = reference_ingest(g)
    nhid, nclass, p = 16, 5, 0.5
    rng = np.random.default_rng(1)
    target = th.tensor(rng.integers(0, nclass, size=n_docs)).long()
    index = th.tensor(np.sort(rng.choice(n_docs, size=200, replace=False))).long()
    out = {
        "n_docs": n_docs, "n_topics": n_topics, "nhid": nhid, "nclass": nclass, "p": p,
        "raw_rows": raw.row.astype(np.int64), "raw_cols": raw.col.astype(np.int64), "raw_vals": raw.data.astype(np.float32),
        "adj_rows": adj._indices()[0].numpy(), "adj_cols": adj._indices()[1].numpy(), "adj_vals": adj._values().numpy(),
        "target": target.numpy(), "index": index.numpy(),
    }
    # ---- featureless ----------------------------------------------------------------------------------------
    th.manual_seed(0)
    model = ref_layer.GCN(nfeat=n, nhid=nhid, nclass=nclass, dropout=p)
    for k, v in model.state_dict().items():
        out["fl_" + k] = v.numpy().copy()
    x = sparse_identity(n)
    ev = run_reference(model, x, adj, target, index, train=False, mseed=7)
    tr = run_reference(model, x, adj, target, index, train=True, mseed=123)
    out["fl_eval_logits"], out["fl_eval_loss"] = ev["logits"], ev["loss"]
    out["fl_train_logits"], out["fl_train_loss"] = tr["logits"], tr["loss"]
    out["fl_keep_mask"] = drawn_mask(123, n, nhid, p)
    for k, v in tr["grads"].items():
        out["fl_grad_" + k] = v
    # ---- sparse features (the reference's real mode: X sparse COO, trainer.py:238) --------------------------------
    nfeat = 24
    rows = np.repeat(np.arange(n), 6)
    cols = np.concatenate([np.sort(rng.choice(nfeat, size=6, replace=False)) for _ in range(n)])
    vals = rng.random(rows.size).astype(np.float32)
    xs = th.sparse_coo_tensor(th.tensor(np.stack([rows, cols])), th.tensor(vals), (n, nfeat))
    th.manual_seed(1)
    model2 = ref_layer.GCN(nfeat=nfeat, nhid=nhid, nclass=nclass, dropout=p)
    for k, v in model2.state_dict().items():
        out["sf_" + k] = v.numpy().copy()
    out["sf_x_rows"], out["sf_x_cols"], out["sf_x_vals"] = rows.astype(np.int64), cols.astype(np.int64), vals
    ev = run_reference(model2, xs, adj, target, index, train=False, mseed=7)
    tr = run_reference(model2, xs, adj, target, index, train=True, mseed=321)
    out["sf_eval_logits"], out["sf_train_logits"], out["sf_train_loss"] = ev["logits"], tr["logits"], tr["loss"]
    out["sf_keep_mask"] = drawn_mask(321, n, nhid, p)
    for k, v in tr["grads"].items():
        out["sf_grad_" + k] = v
    # ---- a lone GraphConvolution and a lone spmm ------------------------------------------------------------------
    th.manual_seed(2)
    gc = ref_layer.GraphConvolution(nfeat, 8)
    out["gc_weight"], out["gc_bias"] = gc.weight.detach().numpy().copy(), gc.bias.detach().numpy().copy()
    out["gc_out"] = gc.forward(xs, adj).detach().numpy().copy()
    B = th.tensor(rng.normal(size=(n, 20)).astype(np.float32))
    out["spmm_B"], out["spmm_Y"] = B.numpy(), th.spmm(adj, B).numpy()
    np.savez_compressed(os.path.join(HERE, "small_dtt.npz"), **out)
    print("small_dtt.npz:", n, "nodes,", adj._nnz(), "nnz")


def load_r8_labels():
    import pandas as pd

    fn = os.path.join(REF, "data/text_dataset/R8.txt")
    df = pd.read_csv(fn, sep="\t", header=None)
    labels = sorted(set(df[2]))  # deterministic id order (the reference uses set() order, trainer.py:254)
    target = np.array([labels.index(v) for v in df[2]], dtype=np.int64)
    train = np.array([i for i, s in enumerate(df[1]) if s in {"train", "training", "20news-bydate-train"}], dtype=np.int64)
    test = np.array([i for i, s in enumerate(df[1]) if s not in {"train", "training", "20news-bydate-train"}], dtype=np.int64)
    return target, train, test, len(labels)


def make_r8(graph_file: str):
    from sklearn.model_selection import train_test_split

    graph = nx.read_weighted_edgelist(graph_file, nodetype=int)  # trainer.py:98
    n = graph.number_of_nodes()
    adj, raw = reference_ingest(graph)
    target_np, train_all, test_lst, nclass = load_r8_labels()
    n_docs = target_np.size
    nhid, p = 200, 0.5
    target = th.tensor(target_np).long()
    x = sparse_identity(n)
    out = {
        "n_docs": n_docs, "n_topics": n - n_docs, "nhid": nhid, "nclass": nclass, "p": p,
        "raw_rows": raw.row.astype(np.int32), "raw_cols": raw.col.astype(np.int32), "raw_vals": raw.data.astype(np.float32),
        "adj_rows": adj._indices()[0].numpy().astype(np.int32), "adj_cols": adj._indices()[1].numpy().astype(np.int32),
        "adj_vals": adj._values().numpy(), "target": target_np.astype(np.int16), "train_all": train_all.astype(np.int32),
        "test": test_lst.astype(np.int32),
    }
    # seed-0 single step (parameters are regenerated from the seed by the tests: same torch call order)
    seed = 0
    train_lst, val_lst = train_test_split(train_all.tolist(), test_size=0.1, shuffle=True, random_state=seed)
    th.manual_seed(seed)
    model = ref_layer.GCN(nfeat=n, nhid=nhid, nclass=nclass, dropout=p)
    index = th.tensor(train_lst).long()
    ev = run_reference(model, x, adj, target, index, train=False, mseed=1)
    tr = run_reference(model, x, adj, target, index, train=True, mseed=mask_seed(seed, 0))
    out["s0_eval_logits"] = ev["logits"]
    out["s0_eval_loss"], out["s0_train_loss"] = ev["loss"], tr["loss"]
    out["s0_train_logits_rows"] = tr["logits"][:: 97].copy()
    out["s0_grad_gc2.weight"], out["s0_grad_gc1.bias"], out["s0_grad_gc2.bias"] = (
        tr["grads"]["gc2.weight"], tr["grads"]["gc1.bias"], tr["grads"]["gc2.bias"])
    out["s0_grad_gc1.weight_rows"] = tr["grads"]["gc1.weight"][:: 97].copy()
    out["s0_grad_gc1.weight_topics"] = tr["grads"]["gc1.weight"][n_docs:].copy()
    np.savez_compressed(os.path.join(HERE, "r8_topic.npz"), **out)
    print("r8_topic.npz:", n, "nodes,", adj._nnz(), "nnz,", nclass, "classes")

    # ---- 5 pinned seeds through the reference training loop (trainer.py:349-398) ------------------------------------
    results = []
    crit = th.nn.CrossEntropyLoss()
    for seed in range(5):
        train_lst, val_lst = train_test_split(train_all.tolist(), test_size=0.1, shuffle=True, random_state=seed)
        th.manual_seed(seed)
        np.random.seed(seed)
        model = ref_layer.GCN(nfeat=n, nhid=nhid, nclass=nclass, dropout=p)
        opt = th.optim.Adam(model.parameters(), lr=0.02)
        tr_i, va_i, te_i = (th.tensor(v).long() for v in (train_lst, val_lst, test_lst))
        stopper = ref_utils.EarlyStopping(10)
        hist = []
        for epoch in range(200):
            model.train()
            opt.zero_grad()
            th.manual_seed(mask_seed(seed, epoch))
            logits = model.forward(x, adj)
            loss = crit(logits[tr_i], target[tr_i])
            loss.backward()
            opt.step()
            model.eval()
            with th.no_grad():
                lg = model.forward(x, adj)
                vloss = float(crit(lg[va_i], target[va_i]).item())
                vacc = ref_utils.accuracy(lg[va_i], target[va_i])
            hist.append({"epoch": epoch, "train_loss": float(loss.item()), "val_loss": vloss, "val_acc": vacc})
            if stopper(vloss):
                break
        model.eval()
        with th.no_grad():
            lg = model.forward(x, adj)
            tacc = ref_utils.accuracy(lg[te_i], target[te_i])
            tloss = float(crit(lg[te_i], target[te_i]).item())
        results.append({"seed": seed, "epochs": len(hist), "test_acc": tacc, "test_loss": tloss, "history": hist})
        print(f"seed {seed}: {len(hist)} epochs, test acc {tacc:.4f}")
    with open(os.path.join(HERE, "r8_training.json"), "w") as fh:
        json.dump({"config": {"nhid": nhid, "dropout": p, "lr": 0.02, "max_epoch": 200, "early_stopping": 10,
                              "val_ratio": 0.1, "featureless": True, "torch": th.__version__},
                   "mask_seed": "1000003*(seed+1)+epoch", "runs": results}, fh)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--r8-graph", default="/tmp/ref_work/data/graph/R8_topic.txt")
    ap.add_argument("--skip-r8", action="store_true")
    args = ap.parse_args()
    make_small()
    if not args.skip_r8 and os.path.exists(args.r8_graph):
        make_r8(args.r8_graph)
