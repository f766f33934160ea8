#!/usr/bin/env python
"""bench.py — TopicGCN graph-convolution hot path on B200: GCN fwd+bwd epochs/s and SpMM GB/s vs the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one train-mode epoch of the 2-layer GCN on the full graph: forward, masked cross-entropy, backward to
all four parameter gradients (reference trainer.py:354-361 without the optimizer; `with_adam` is reported next to
it).  Workload at N=1: BASELINE.json configs[2], the 1M-document x 256-topic synthetic graph (featureless X = I,
hidden 256, 20 classes) — the largest single-GPU configuration and the one whose operands exceed the 126 MB L2.
At N>1 every rank holds a shard of that shape (documents row-sharded, topics replicated, see shard.py): weak scaling;
`value` is epochs/s multiplied by the number of 1M-document shards, i.e. whole-job throughput in C3-equivalents.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the reference path (all host
threads) on a bounded slice of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c3": "c3_1m_docs_256_topics",
    "c2": "c2_20ng_shape",
    "c4": "c4_shard_6p25m_docs_1024_topics",
    "c5": "c5_textgcn_r8_shape",
    "c1": "c1_r8_shape",
}
METRIC = "gcn_fwd_bwd_epochs_per_sec"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.lines, self.gpu = None, [], gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# per-kernel CUDA-event hook
# ---------------------------------------------------------------------------------------------------------------
class EventHook:
    """Records a CUDA event pair on the launching (current torch) stream around selected C-ABI calls."""

    def __init__(self, tags):
        self.tags, self.records, self.enabled = set(tags), [], False

    def start(self, tag, info):
        import torch
        if not self.enabled or tag not in self.tags:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        return (tag, info, a, b)

    def stop(self, tok):
        tag, info, a, b = tok
        b.record()
        self.records.append((tag, info.get("n_feat"), info.get("csr"), a, b))

    def summary(self):
        out = {}
        for tag, f, csr, a, b in self.records:
            out.setdefault((tag, f, id(csr)), [csr, []])[1].append(a.elapsed_time(b))
        return out


def spmm_bytes(n_rows: int, n_cols: int, nnz: int, n_feat: int) -> float:
    """Algorithmic bytes of Y[n_rows x F] = A B: idx+val stream, row pointers, each B row once, Y once."""
    return nnz * 8.0 + (n_rows + 1) * 4.0 + n_cols * n_feat * 4.0 + n_rows * n_feat * 4.0


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline of our arm, and the whole --impl reference arm)
# ---------------------------------------------------------------------------------------------------------------
def cpu_epoch_rate(workload: str, sample_docs: int, steps: int, warmup: int):
    """epochs/s of the oracle port (numpy + OpenMP C loop, all host threads) on a document slice of the workload;
    returns (epochs/s scaled to the full workload, description, threads, per-step ms on the slice)."""
    import numpy as np
    import torch
    from oracle import gcn_oracle as O
    from topicgcn_b200 import graphgen

    name = WORKLOADS[workload]
    full_docs = graphgen.CONFIGS[name][1].get("n_docs", 7674)
    sample_docs = min(sample_docs, full_docs)
    g, hidden, n_class = graphgen.make_config(name, device="cpu", scale=sample_docs / full_docs)
    threads = os.cpu_count() or 1
    O.set_threads(threads)
    torch.set_num_threads(threads)
    coo = O.Coo(g.rows.numpy(), g.cols.numpy(), g.vals.numpy(), (g.n, g.n))
    rng = np.random.default_rng(0)
    sd = 1.0 / np.sqrt(hidden)
    params = {"gc1.weight": rng.uniform(-sd, sd, size=(g.n, hidden)).astype(np.float32),
              "gc1.bias": rng.uniform(-sd, sd, size=hidden).astype(np.float32),
              "gc2.weight": rng.uniform(-0.2, 0.2, size=(hidden, n_class)).astype(np.float32),
              "gc2.bias": rng.uniform(-0.2, 0.2, size=n_class).astype(np.float32)}
    target, index = g.labels.numpy(), g.train_idx.numpy()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        # the reference draws its Bernoulli mask inside the step (th.dropout, layer.py:185): same torch call, timed
        mask = torch.empty(g.n, hidden).bernoulli_(0.5).numpy()
        O.gcn_loss_and_grads(None, coo, params, target, index, p=0.5, training=True, keep_mask=mask)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    O.set_threads(1)
    per_step = statistics.median(times)
    frac = g.n_docs / full_docs
    rate_full = (1.0 / per_step) * frac
    desc = (f"{g.n_docs} of {full_docs} documents ({frac:.3g} of the workload, same topics/hidden/classes), "
            f"{steps} fwd+bwd epochs after {warmup} warm-up; epochs/s scaled by {frac:.3g} to the full workload")
    return rate_full, desc, threads, per_step * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, desc, threads, ms = cpu_epoch_rate(args.workload, args.cpu_sample_docs, max(1, min(args.steps, 5)),
                                             max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "epochs/s", "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 5)), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "featureless": True, "optimizer_in_step": False},
        "cpu_baseline": {"value": rate, "unit": "epochs/s", "cores": threads, "kind": "port", "sample": desc,
                         "ms_per_step_on_sample": ms},
        "e2e": {"value": rate, "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import topicgcn_b200 as tg
    from topicgcn_b200 import graphgen, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line of the contract: NCCL's version/debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    tg._native.lib()

    name = WORKLOADS[args.workload]
    hidden, n_class = graphgen.CONFIGS[name][2], graphgen.CONFIGS[name][3]
    hook = EventHook({"spmm", "gc1_fwd", "gc2_loss_fwd", "stream_spmm"})
    ops.set_kernel_hook(hook)

    if world == 1:
        g, hidden, n_class = graphgen.make_config(name, device=dev, seed=0)
        adj = g.adj()
        n = g.n
        torch.manual_seed(0)
        model = tg.GCN(n, hidden, n_class, 0.5).to(dev)
        model.train()
        x = tg.Featureless(n)
        csr = tg.cached_csr(adj)
        csr.transpose()  # symmetry check / transposed copy is plan-time work
        row_label = ops.make_row_label(n, g.labels, g.train_idx)
        labels_host = g.labels.cpu().pin_memory()
        index_host = g.train_idx.cpu().pin_memory()
        n_docs_total, nnz_total = g.n_docs, g.nnz

        def step_device():
            for p in model.parameters():
                p.grad = None
            loss = model.loss(x, adj, g.labels, g.train_idx, row_label=row_label)
            loss.backward()
            return loss

        def step_e2e():
            # reference-facing call with HOST inputs: labels + train index come from pinned host memory every
            # step, the loss goes back to the host (trainer.py:357-367: forward, loss, backward, loss.item())
            for p in model.parameters():
                p.grad = None
            loss = model.loss(x, adj, labels_host, index_host)  # host tensors: copied on a side stream inside the call
            loss.backward()
            return float(loss.item())

        h2d = labels_host.numel() * 8 + index_host.numel() * 8
        params = list(model.parameters())
        shard_desc = "single GPU, no collective"
    else:
        from topicgcn_b200 import shard
        sg = shard.make_sharded_config(name, rank=rank, world=world, device=dev, seed=0)
        torch.manual_seed(rank)
        model = shard.ShardedGCN(sg, hidden, n_class, 0.5).to(dev)
        model.sync_replicated()
        model.train()
        csr = shard._csr(sg)
        labels_host = sg.labels.cpu().pin_memory()
        index_host = sg.train_idx.cpu().pin_memory()
        row_label = ops.make_row_label(sg.n_local, sg.labels, sg.train_idx)
        n_docs_total, nnz_total = sg.n_docs_global, sg.nnz_global

        def step_device():
            for p in model.parameters():
                p.grad = None
            loss = model.loss(row_label=row_label)
            loss.backward()
            return loss

        def step_e2e():
            for p in model.parameters():
                p.grad = None
            loss = model.loss(labels=labels_host, index=index_host)  # host tensors: copied on a side stream inside the call
            loss.backward()
            return float(loss.item())

        h2d = labels_host.numel() * 8 + index_host.numel() * 8
        params = list(model.parameters())
        shard_desc = f"documents row-sharded over {world} ranks, topic rows replicated, NCCL all-reduce of K x F partials"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, use_hook=False):
        for _ in range(warmup):
            fn()
        barrier()
        ops.Stats.launches = 0
        hook.enabled = use_hook
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        hook.enabled = False
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.Stats.launches, clocks

    # ---- device-resident timing (value) + live per-kernel events ------------------------------------------------------
    ms_total, launches, clocks = timed(step_device, args.steps, args.warmup, use_hook=True)
    ms_step = ms_total / args.steps
    # ---- with Adam (reported, not the headline) -----------------------------------------------------------------------
    opt = tg.optim.Adam(params, lr=0.02)  # one tg_adam_f32 pass per parameter (torch.optim.Adam semantics, trainer.py:307)

    def step_adam():
        step_device()
        opt.step()

    ms_adam, _, _ = timed(step_adam, max(3, args.steps // 2), 3)
    ms_adam /= max(3, args.steps // 2)
    del opt
    # ---- the same step replayed from a CUDA graph (launch-bound small graphs; reported, not the headline) -----------
    graph_info = None
    if world == 1:
        try:
            cap = tg.CapturedTrainStep(model, x, adj, g.labels, g.train_idx)
            ms_g, _, _ = timed(cap.step, args.steps, args.warmup)
            graph_info = {"ms_per_step": ms_g / args.steps, "value": 1e3 * args.steps / ms_g}
            del cap
            model._offset_dev = None
        except Exception as exc:  # pragma: no cover
            graph_info = {"error": str(exc)[:200]}
    # ---- end to end through the public API with host inputs -----------------------------------------------------------
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
    ms_e2e /= args.steps

    shards = world  # one C3-shaped shard per rank
    value = shards * 1e3 / ms_step
    e2e_value = shards * 1e3 / ms_e2e

    # ---- roofline of the dominant kernel: the F = hidden SpMM (layer-1 forward and the dW1 backward) ----------------
    hbm_peak, peak_src = measured_peaks()
    kern = {}
    for (tag, f, _cid), (kcsr, ts) in hook.summary().items():
        if f is None or kcsr is None or kcsr.n_rows != csr.n_rows:
            continue  # (the sharded mode also runs the epilogue on the K replicated topic rows: not the SpMM)
        kern.setdefault((tag, f), []).append((kcsr, ts))
    roof = None
    detail = {}
    for (tag, f), lst in kern.items():
        ts = [t for _, tl in lst for t in tl]
        kcsr = lst[0][0]
        nbytes = spmm_bytes(kcsr.n_rows, kcsr.n_cols, kcsr.nnz, f)
        avg_ms = sum(ts) / len(ts)
        detail[f"{tag}_F{f}"] = {"launches": len(ts), "avg_ms": avg_ms, "algorithmic_GB": nbytes / 1e9,
                                 "GBps": nbytes / 1e6 / avg_ms, "frac_of_peak": nbytes / 1e6 / avg_ms / hbm_peak}
    dom = [(k, v) for k, v in detail.items() if k.endswith(f"_F{hidden}")]
    if dom:
        tot_ms = sum(v["avg_ms"] * v["launches"] for _, v in dom)
        tot_bytes = sum(v["algorithmic_GB"] * v["launches"] for _, v in dom)
        n_l = sum(v["launches"] for _, v in dom)
        achieved = tot_bytes * 1e3 / tot_ms
        if ops.roles2_path(csr, hidden) == "wide":
            kname = (f"roles2_kernel + stream_finish_kernel (warp-per-slot role-specialised column-chunk streaming SpMM: "
                     f"TMA-staged chunks, bulk-copy entry ring, FFMA2; F={hidden})")
        elif csr.streaming:
            kname = (f"stream_roles_kernel<512,4,KPG,EpiStore> + stream_finish_kernel (role-specialised column-chunk "
                     f"streaming SpMM with TMA-staged chunks, F={hidden})")
        else:
            kname = f"spmm_kernel<4,32,{(hidden // 4 + 31) // 32},EpiStore> (gather SpMM, F={hidden})"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if csr.streaming and world == 1 and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(name)  # dram bytes per launch from the committed ncu --set full capture
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "kernel": kname,
                "launches_timed": n_l, "avg_launch_ms": tot_ms / n_l, "algorithmic_bytes_per_launch": tot_bytes * 1e9 / n_l,
                "share_of_step": tot_ms / (ms_step * args.steps), "peak_source": peak_src,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "bytes_formula": "nnz*8 + (n_rows+1)*4 + n_cols*F*4 + n_rows*F*4 (SURVEY 8d)"}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, desc, threads, ms = cpu_epoch_rate(args.workload, args.cpu_sample_docs, 3, 1)
            cpu = {"value": rate, "unit": "epochs/s", "cores": threads, "kind": "port", "sample": desc,
                   "ms_per_step_on_sample": ms}
        line = {
            "metric": METRIC, "value": value, "unit": "epochs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "docs_total": n_docs_total, "nnz_total": nnz_total, "hidden": hidden,
                       "classes": n_class, "featureless": True, "optimizer_in_step": False,
                       "l2_policy": "inputs larger than L2 (W1 alone is %.2f GB per GPU)" % (csr.n_cols * hidden * 4 / 1e9),
                       "parallelism": shard_desc, "value_definition": "epochs/s x number of per-GPU shards"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "epochs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "roofline": roof, "kernels": detail, "cuda_graph_step": graph_info, "with_adam": {"ms_per_step": ms_adam, "value": shards * 1e3 / ms_adam},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-docs", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
