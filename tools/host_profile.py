#!/usr/bin/env python
"""Host-side latency of the fused train step up to its first kernel launch (GPU box): what a synchronous loop (loss.item()
every step) exposes once per step.  python tools/host_profile.py [--workload c3]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen, layer, ops  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g, hidden, n_class = graphgen.make_config(WORKLOADS[a.workload], device=dev)
    n, adj = g.n, g.adj()
    model = tg.GCN(n, hidden, n_class, 0.5).to(dev)
    model.train()
    x = tg.Featureless(n)
    labels_host, index_host = g.labels.cpu().pin_memory(), g.train_idx.cpu().pin_memory()

    def step():
        for p in model.parameters():
            p.grad = None
        loss = model.loss(x, adj, labels_host, index_host)
        loss.backward()
        return float(loss.item())

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pc = time.perf_counter
    rows = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = pc()
        csr = layer._as_csr(adj)
        t1 = pc()
        S1 = layer._support(x, model.gc1.weight)
        mask, seed, off = model._dropout_state()
        t2 = pc()
        H1 = ops.gc1_forward(csr, S1.detach(), model.gc1.bias.detach(), 0.5, True, mask, seed, off)
        t3 = pc()
        S2 = ops.dense_nn(H1, model.gc2.weight.detach())
        t4 = pc()
        rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    rows = rows[5:]
    med = [sorted(r[i] for r in rows)[len(rows) // 2] * 1e6 for i in range(4)]
    print("host microseconds (median): cached_csr %.1f, support + dropout state %.1f, gc1_forward (3 launches) %.1f, dense_nn %.1f" % tuple(med))
    # the C call of gc1 alone
    ws, ws_bytes = csr.workspace(hidden)
    from topicgcn_b200 import _native as N
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = pc()
        ops.spmm(csr, S1.detach(), None, out=H1)
        ts.append(pc() - t0)
    print("host microseconds (median): spmm wrapper + C call (tensor-map encodes, 2 launches) %.1f" % (sorted(ts)[10] * 1e6))
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = pc()
        torch.empty((n, hidden), dtype=torch.float32, device=dev)
        ts.append(pc() - t0)
    print("host microseconds (median): torch.empty of the 1 GB output %.1f" % (sorted(ts)[10] * 1e6))


if __name__ == "__main__":
    main()
