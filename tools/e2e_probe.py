#!/usr/bin/env python
"""Where the end-to-end step (host labels in, loss out, one sync per step) loses time against the device-timed step (GPU box):
    python tools/e2e_probe.py [--workload c3 --steps 20]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen, ops  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g, hidden, n_class = graphgen.make_config(WORKLOADS[a.workload], device=dev)
    n = g.n
    adj = g.adj()
    torch.manual_seed(0)
    model = tg.GCN(n, hidden, n_class, 0.5).to(dev)
    model.train()
    x = tg.Featureless(n)
    csr = tg.cached_csr(adj)
    csr.transpose()
    row_label = ops.make_row_label(n, g.labels, g.train_idx)
    labels_host = g.labels.cpu().pin_memory()
    index_host = g.train_idx.cpu().pin_memory()

    def zero():
        for p in model.parameters():
            p.grad = None

    def dev_async():
        zero()
        model.loss(x, adj, g.labels, g.train_idx, row_label=row_label).backward()

    def dev_sync():
        zero()
        loss = model.loss(x, adj, g.labels, g.train_idx, row_label=row_label)
        loss.backward()
        return float(loss.item())

    def dev_labels_sync():
        zero()
        loss = model.loss(x, adj, g.labels, g.train_idx)
        loss.backward()
        return float(loss.item())

    def e2e():
        zero()
        loss = model.loss(x, adj, labels_host, index_host)
        loss.backward()
        return float(loss.item())

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / a.steps

    for name, fn in (("device labels, no sync", dev_async), ("device row_label, loss.item() per step", dev_sync),
                     ("device labels -> row_label per step, loss.item()", dev_labels_sync), ("host labels, loss.item() (e2e)", e2e)):
        print(f"{name:55s} {timed(fn):.3f} ms")
    # host time of one step's Python + launches when nothing waits (GPU idle at the start)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    zero()
    loss = model.loss(x, adj, g.labels, g.train_idx, row_label=row_label)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host time to queue the forward {1e3 * (t1 - t0):.3f} ms, the backward {1e3 * (t2 - t1):.3f} ms")


if __name__ == "__main__":
    main()
