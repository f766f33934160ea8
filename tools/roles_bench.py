#!/usr/bin/env python
"""Per-role timing of the role-specialised streaming SpMM on a named workload (GPU box).

    python tools/roles_bench.py [--workload c4] [--feat 256] [--reps 5]
TG_ROLES_ONLY (1 = hub role alone, 2 = document role alone) and the other TG_ROLES2_* knobs are read when a plan is
created, so every configuration below builds its own plan from the same CSR arrays."""
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
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--feat", default="256")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=None)
    ap.add_argument("--configs", default="both;TG_ROLES_ONLY=1;TG_ROLES_ONLY=2")
    ap.add_argument("--op", default="spmm", choices=["spmm", "loss"], help="plain product or the fused loss forward (class-sized --feat)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g, hidden, n_class = graphgen.make_config(WORKLOADS[args.workload], device=dev, scale=args.scale)
    peak, _ = measured_peaks()
    base = tg.DeviceCSR.from_coo(g.rows, g.cols, g.vals, g.n, g.n)
    print("plan:", dict(hub_rows=base.n_hub_rows, gs=base.hub_gs, groups=base.hub_groups, nq=base.doc_nq, T=base.chunk_rows), flush=True)
    out = {}
    for cfg in args.configs.split(";"):
        env = dict(kv.split("=") for kv in cfg.split(",") if "=" in kv)
        for k, v in env.items():
            os.environ[k] = v
        csr = tg.DeviceCSR(base.rowptr, base.colidx, base.vals, g.n, g.n)
        for k in env:
            del os.environ[k]
        for F in [int(f) for f in args.feat.split(",")]:
            B = torch.randn(g.n, F, device=dev)
            Y = torch.empty(g.n, F, device=dev)
            bias = torch.randn(F, device=dev)
            row_label = ops.make_row_label(g.n, g.labels % F, g.train_idx)

            def run():
                if args.op == "spmm":
                    ops.spmm(csr, B, None, out=Y)
                else:
                    ops.gc2_loss_forward(csr, B, bias, row_label, 1.0 / g.train_idx.numel(), want_logits=False, want_grad=True)

            for _ in range(2):
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
            out[f"{cfg}|F{F}"] = {"ms": round(ms, 4), "frac": round(nbytes / 1e6 / ms / peak, 4)}
            print(cfg, F, out[f"{cfg}|F{F}"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
