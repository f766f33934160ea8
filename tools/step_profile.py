#!/usr/bin/env python
"""Kernel-level breakdown of one train step (torch.profiler / CUPTI; no nsys in the image).

    python tools/step_profile.py [--workload c3] [--steps 5]                      # single GPU
    torchrun --nproc-per-node 2 tools/step_profile.py --steps 5                   # sharded step, rank 0 reports
Prints, per kernel name, launches per step and device time per step, plus the span of a step on the device (first kernel
start to last kernel end, averaged) so that gaps show up as span - sum."""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import topicgcn_b200 as tg  # noqa: E402
from topicgcn_b200 import graphgen, ops  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    name = WORKLOADS[args.workload]
    hidden, n_class = graphgen.CONFIGS[name][2], graphgen.CONFIGS[name][3]
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from topicgcn_b200 import shard
        sg = shard.make_sharded_config(name, rank=rank, world=world, device=dev, seed=0)
        model = shard.ShardedGCN(sg, hidden, n_class, 0.5).to(dev)
        model.sync_replicated()
        model.train()
        row_label = ops.make_row_label(sg.n_local, sg.labels, sg.train_idx)

        def step():
            for p in model.parameters():
                p.grad = None
            model.loss(row_label=row_label).backward()
    else:
        g, hidden, n_class = graphgen.make_config(name, device=dev, seed=0)
        adj = g.adj()
        model = tg.GCN(g.n, hidden, n_class, 0.5).to(dev)
        model.train()
        x = tg.Featureless(g.n)
        row_label = ops.make_row_label(g.n, g.labels, g.train_idx)

        def step():
            for p in model.parameters():
                p.grad = None
            model.loss(x, adj, g.labels, g.train_idx, row_label=row_label).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        agg = collections.OrderedDict()
        for e in evs:
            a = agg.setdefault(e.name[:90], [0, 0.0])
            a[0] += 1
            a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        t0 = min(e.time_range.start for e in evs)
        t1 = max(e.time_range.end for e in evs)
        total = sum(v[1] for v in agg.values())
        print(f"device span per step {(t1 - t0) / args.steps / 1e3:.3f} ms, sum of kernel times per step {total / args.steps / 1e3:.3f} ms "
              f"({world} rank(s), {name})")
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
        for k, (n, t) in rows:
            print(f"{t / args.steps / 1e3:8.4f} ms  x{n / args.steps:5.1f}  {k}")
        if args.out:
            json.dump({"world": world, "workload": name, "span_ms_per_step": (t1 - t0) / args.steps / 1e3,
                       "kernels": [{"name": k, "launches_per_step": n / args.steps, "ms_per_step": t / args.steps / 1e3} for k, (n, t) in rows]},
                      open(args.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
