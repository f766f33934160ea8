#!/usr/bin/env python
"""Micro-benchmark of the skinny dense kernels (GPU box): python tools/dense_bench.py [--n 1000256 --h 256 --c 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topicgcn_b200 import ops  # noqa: E402


def timeit(fn, reps=20):
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
    ap.add_argument("--n", type=int, default=1000256)
    ap.add_argument("--h", type=int, default=256)
    ap.add_argument("--c", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    # H1 as the layer produces it: relu then dropout(0.5) -> ~75 % exact zeros
    H1 = torch.relu(torch.randn(a.n, a.h, device=dev)) * (torch.rand(a.n, a.h, device=dev) < 0.5) * 2.0
    W2 = torch.randn(a.h, a.c, device=dev)
    dS2 = torch.randn(a.n, a.c, device=dev)
    X = torch.randn(a.n, a.c, device=dev)
    gflop = 2.0 * a.n * a.h * a.c / 1e9
    ms = timeit(lambda: ops.dense_nn(H1, W2))
    print(f"dense_nn      {ms:.3f} ms  {gflop / ms:.1f} TFLOP/s  {a.n * a.h * 4 / 1e6 / ms:.0f} GB/s (A read)")
    ms = timeit(lambda: ops.hidden_backward(H1, dS2, W2, 2.0))
    print(f"hidden_bwd    {ms:.3f} ms  {2 * gflop / ms:.1f} TFLOP/s  {2 * a.n * a.h * 4 / 1e6 / ms:.0f} GB/s (H1 read + dZ1 write)")
    ms = timeit(lambda: torch.mm(H1, W2))
    print(f"cuBLAS mm     {ms:.3f} ms (for scale)")
    ms = timeit(lambda: ops.colsum(X))
    print(f"colsum        {ms:.3f} ms")


if __name__ == "__main__":
    main()
