#!/usr/bin/env python
"""F = hidden product alone on a named workload (GPU box; short enough for an ncu capture):
    [TG_ROLES_ONLY=1|2] python tools/wide_bench.py [--workload c4 --scale 0.3 --feat 256 --reps 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen, ops  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--scale", type=float, default=0.3)
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g, _, _ = graphgen.make_config(WORKLOADS[a.workload], device=dev, scale=a.scale)
    csr = tg.DeviceCSR.from_coo(g.rows, g.cols, g.vals, g.n, g.n)
    B = torch.randn(g.n, a.feat, device=dev)
    Y = torch.empty(g.n, a.feat, device=dev)
    for _ in range(2):
        ops.spmm(csr, B, None, out=Y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        ops.spmm(csr, B, None, out=Y)
    e1.record()
    torch.cuda.synchronize()
    print(f"spmm {a.workload} x{a.scale} F={a.feat} {e0.elapsed_time(e1) / a.reps:.4f} ms (TG_ROLES_ONLY={os.environ.get('TG_ROLES_ONLY', '0')})")


if __name__ == "__main__":
    main()
