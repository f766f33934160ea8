#!/usr/bin/env python
"""Hidden-layer backward alone at the C3 shape (GPU box; short enough for an ncu capture):
    python tools/hidden_bench.py [--n 1000256 --h 256 --c 20 --reps 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from topicgcn_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000256)
    ap.add_argument("--h", type=int, default=256)
    ap.add_argument("--c", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    H1 = torch.relu(torch.randn(a.n, a.h, device=dev)) * (torch.rand(a.n, a.h, device=dev) < 0.5) * 2.0
    W2 = torch.randn(a.h, a.c, device=dev)
    dS2 = torch.randn(a.n, a.c, device=dev)
    out = torch.empty_like(H1)
    for _ in range(3):
        ops.hidden_backward(H1, dS2, W2, 2.0, out_dZ1=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        ops.hidden_backward(H1, dS2, W2, 2.0, out_dZ1=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(f"hidden_bwd {ms:.3f} ms  {2 * a.n * a.h * 4 / 1e6 / ms:.0f} GB/s (H1 read + dZ1 write)")


if __name__ == "__main__":
    main()
