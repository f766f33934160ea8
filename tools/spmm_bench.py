#!/usr/bin/env python
"""Micro-benchmark of the SpMM kernels on a named synthetic workload (GPU box).

    python tools/spmm_bench.py [--workload c3] [--feat 256,20] [--reps 20]
Env knobs of the library (TG_ROLES2_*, read at plan creation) apply.  Prints GB/s against the algorithmic bytes.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen, ops  # noqa: E402
from bench import WORKLOADS, measured_peaks, spmm_bytes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--feat", default="256,20")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--scale", type=float, default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g, hidden, n_class = graphgen.make_config(WORKLOADS[args.workload], device=dev, scale=args.scale)
    peak, _ = measured_peaks()
    out = {"workload": WORKLOADS[args.workload], "n": g.n, "nnz": g.nnz, "env": {k: v for k, v in os.environ.items() if k.startswith("TG_")}}
    for streaming in (True, False):
        csr = tg.DeviceCSR.from_coo(g.rows, g.cols, g.vals, g.n, g.n, streaming=streaming)
        for F in [int(f) for f in args.feat.split(",")]:
            B = torch.randn(g.n, F, device=dev)
            bias = torch.randn(F, device=dev)
            Y = torch.empty(g.n, F, device=dev)
            for mode in ("spmm", "gc1"):
                def run():
                    if mode == "spmm":
                        ops.spmm(csr, B, None, out=Y)
                    else:
                        ops.gc1_forward(csr, B, bias, 0.5, True, seed=1, offset=2, out=Y)
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                nbytes = spmm_bytes(g.n, g.n, g.nnz, F)
                key = f"{'stream' if csr.streaming else 'gather'}_{mode}_F{F}"
                out[key] = {"ms": round(ms, 4), "GBps": round(nbytes / 1e6 / ms, 1), "frac": round(nbytes / 1e6 / ms / peak, 4)}
                print(key, out[key], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
