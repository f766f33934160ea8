#!/usr/bin/env python
"""Time one train step in the reference's REAL feature mode (sparse X = L2-normalised [theta | topic embedding],
trainer.py:197-238) next to the featureless mode, on a synthetic graph of the C3 shape (GPU box).
    python tools/feature_mode_bench.py [--docs 1000000] [--topics 256] [--emb 100]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--topics", type=int, default=256)
    ap.add_argument("--emb", type=int, default=100)
    ap.add_argument("--theta-nnz", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = graphgen.doc_topic_topic_graph(a.docs, a.topics, 8, 8, True, 20, 0, dev)
    nfeat = max(a.topics, a.emb)
    gen = torch.Generator(device=dev).manual_seed(1)
    # documents: theta over `theta_nnz` topics; topic nodes: a dense embedding row (trainer.py:197-238), rows L2-normalised
    cols_d = torch.rand(a.docs, a.topics, device=dev, generator=gen).topk(a.theta_nnz, dim=1).indices.sort(dim=1).values
    vals_d = torch.rand(a.docs, a.theta_nnz, device=dev, generator=gen) + 0.01
    vals_d = vals_d / vals_d.norm(dim=1, keepdim=True)
    rows_d = torch.arange(a.docs, device=dev).unsqueeze(1).expand(-1, a.theta_nnz)
    vals_t = torch.randn(a.topics, a.emb, device=dev, generator=gen)
    vals_t = vals_t / vals_t.norm(dim=1, keepdim=True)
    rows_t = (torch.arange(a.topics, device=dev) + a.docs).unsqueeze(1).expand(-1, a.emb)
    cols_t = torch.arange(a.emb, device=dev).unsqueeze(0).expand(a.topics, -1)
    X = torch.sparse_coo_tensor(torch.stack([torch.cat([rows_d.reshape(-1), rows_t.reshape(-1)]),
                                             torch.cat([cols_d.reshape(-1), cols_t.reshape(-1)])]),
                                torch.cat([vals_d.reshape(-1), vals_t.reshape(-1)]), (g.n, nfeat)).coalesce()
    adj = g.adj()
    for name, x, nf in (("featureless", None, g.n), ("topic features", X, nfeat)):
        torch.manual_seed(0)
        model = tg.GCN(nf, 256, 20, 0.5).to(dev).train()
        row_label = tg.ops.make_row_label(g.n, g.labels, g.train_idx)

        def step():
            for p in model.parameters():
                p.grad = None
            model.loss(x, adj, g.labels, g.train_idx, row_label=row_label).backward()
        ms = timeit(step)
        print(f"{name:15s} nfeat={nf:8d}  {ms:8.3f} ms per fwd+bwd step", flush=True)


if __name__ == "__main__":
    main()
