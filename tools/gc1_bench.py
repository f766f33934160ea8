#!/usr/bin/env python
"""Fused layer-1 forward (SpMM + bias + ReLU + Philox dropout) next to the plain product on a named workload (GPU box):
    python tools/gc1_bench.py [--workload c3 --scale 1.0 --feat 256 --reps 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen, ops  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def timeit(fn, reps):
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
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g, _, _ = graphgen.make_config(WORKLOADS[a.workload], device=dev, scale=a.scale)
    csr = tg.DeviceCSR.from_coo(g.rows, g.cols, g.vals, g.n, g.n)
    B = torch.randn(g.n, a.feat, device=dev)
    Y = torch.empty(g.n, a.feat, device=dev)
    bias = torch.randn(a.feat, device=dev)
    print(f"spmm     {timeit(lambda: ops.spmm(csr, B, None, out=Y), a.reps):.4f} ms")
    print(f"spmm+b   {timeit(lambda: ops.spmm(csr, B, bias, out=Y), a.reps):.4f} ms")
    print(f"gc1 eval {timeit(lambda: ops.gc1_forward(csr, B, bias, 0.5, False, out=Y), a.reps):.4f} ms")
    print(f"gc1 drop {timeit(lambda: ops.gc1_forward(csr, B, bias, 0.5, True, seed=1, offset=0, out=Y), a.reps):.4f} ms")


if __name__ == "__main__":
    main()
