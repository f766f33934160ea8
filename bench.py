#!/usr/bin/env python
"""bench.py — TopicGCN graph-convolution hot path on B200: GCN fwd+bwd epochs/s and SpMM GB/s vs the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one train-mode epoch of the 2-layer GCN on the full graph: forward, masked cross-entropy, backward to
all four parameter gradients (reference trainer.py:354-361 without the optimizer; `with_adam` is reported next to
it).  Workload at N=1: BASELINE.json configs[2], the 1M-document x 256-topic synthetic graph (featureless X = I,
hidden 256, 20 classes) — the largest single-GPU configuration and the one whose operands exceed the 126 MB L2.
At N>1 every rank holds a shard of that shape (documents row-sharded, topics replicated, see shard.py): weak scaling;
`value` is epochs/s multiplied by the number of 1M-document shards, i.e. whole-job throughput in C3-equivalents.  The
N>1 line also carries a `c4` block: the same step on BASELINE.json configs[3] (6.25 M documents x 1 024 topics per
GPU — 50 M documents at 8 GPUs) with its own per-kernel table, and a `consistency` block (replicated gradients bit-identical
on all ranks; sharded loss against a single-GPU run of the same global graph at a small size).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference path on a bounded slice of
the same workload: the OpenMP port (all host threads; the value) and, beside it, the ATen operators the reference
itself runs on (oracle/torch_ref.py: th.spmm on a COO tensor + autograd), which anchors the port.
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
        self.tags, self.records, self.dense, self.enabled = set(tags), [], [], False

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
        if "csr" in info:
            self.records.append((tag, info.get("n_feat"), info.get("csr"), a, b))
        else:
            self.dense.append((tag, info.get("n"), info.get("h"), info.get("c"), a, b))

    def dense_summary(self):
        """{(tag, n, h, c): [ms, ...]} of the skinny dense kernels (H1 W2 and the fused hidden-layer backward)."""
        out = {}
        for tag, n, h, c, a, b in self.dense:
            out.setdefault((tag, n, h, c), []).append(a.elapsed_time(b))
        return out

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


def aten_cpu_step_ms(workload: str, sample_docs: int, steps: int = 1, warmup: int = 1):
    """ms per fwd+bwd step of oracle/torch_ref.py on the CPU — the ATen operators the reference runs on (th.spmm on COO
    tensors, serial in ATen; dense products and elementwise ops on all torch threads) — on the same document slice."""
    import torch
    from oracle import torch_ref as TR
    from topicgcn_b200 import graphgen

    name = WORKLOADS[workload]
    full_docs = graphgen.CONFIGS[name][1].get("n_docs", 7674)
    g, hidden, n_class = graphgen.make_config(name, device="cpu", scale=min(sample_docs, full_docs) / full_docs)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = TR.GCNRef(g.n, hidden, n_class, 0.5)
    model.train()
    x, adj = TR.sparse_identity(g.n, "cpu"), g.adj()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        TR.train_step(model, x, adj, g.labels, g.train_idx)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return statistics.median(times) * 1e3, g.n_docs, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps / --warmup are honoured as given, for BOTH CPU legs; every step is one fwd+bwd epoch on the bounded document
    # slice (1.5 - 2 s per step at the default 200 K documents of C3 on 16 cores):
    #   aten_operators : oracle/torch_ref.py — the reference module's own operators (th.spmm on COO tensors, relu, dropout,
    #                    cross_entropy, autograd) as torch 2.11 runs them on all host threads
    #   openmp_port    : oracle/gcn_oracle.py + spmm_oracle.c — the numpy / OpenMP restatement used as the parity checker
    # The line's value is the FASTER leg (the stronger baseline).
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    rate_p, desc_p, threads, ms_p = cpu_epoch_rate(args.workload, args.cpu_sample_docs, steps, warmup)
    legs = {"openmp_port": {"ms_per_step_on_sample": ms_p, "threads": threads}}
    ms, kind_note = ms_p, "openmp_port"
    try:
        ms_a, docs_a, threads_a = aten_cpu_step_ms(args.workload, args.cpu_sample_docs, steps, warmup)
        legs["aten_operators"] = {"ms_per_step_on_sample": ms_a, "threads": threads_a}
        if ms_a < ms_p:
            ms, kind_note = ms_a, "aten_operators"
    except Exception as exc:  # pragma: no cover
        legs["aten_operators"] = {"error": str(exc)[:200]}
    rate = rate_p * ms_p / ms  # same slice, same scaling to the full workload
    desc = desc_p + f"; value from the faster CPU leg: {kind_note}"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "epochs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup,
        # the time of ONE timed step as it ran (the slice), so that steps x ms_per_step is this leg's own duration;
        # `value` is that rate scaled to the full workload (epochs/s x slice fraction), as `cpu_baseline.sample` says
        "ms_per_step": ms, "ms_per_step_full_workload_extrapolated": 1e3 / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "featureless": True, "optimizer_in_step": False,
                   "sample_docs": args.cpu_sample_docs},
        "cpu_baseline": {"value": rate, "unit": "epochs/s", "cores": threads, "kind": "port", "sample": desc,
                         "ms_per_step_on_sample": ms, "value_from": kind_note, "legs": legs},
        "e2e": {"value": rate, "unit": "epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_library_baseline(g, hidden: int, n_class: int, dev, steps: int = 5, warmup: int = 2):
    """The unmodified library path on the same GPU and the same inputs (SURVEY §2.1's bar): oracle/torch_ref.py on `cuda`
    — torch's COO tensors, per-call coalesce() + cuSPARSE SpMM, unfused elementwise kernels, autograd — timed with CUDA
    events; plus the lone F = hidden product th.spmm(adj, B)."""
    import torch
    from oracle import torch_ref as TR

    torch.manual_seed(0)
    model = TR.GCNRef(g.n, hidden, n_class, 0.5).to(dev)
    model.train()
    x, adj = TR.sparse_identity(g.n, dev), g.adj()

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    ms_step = timed(lambda: TR.train_step(model, x, adj, g.labels, g.train_idx), steps, warmup)
    B = torch.randn(g.n, hidden, device=dev)
    ms_spmm = timed(lambda: torch.spmm(adj, B), steps, warmup)
    adj_c = adj.coalesce()
    ms_spmm_c = timed(lambda: torch.spmm(adj_c, B), steps, warmup)
    del model, x, B, adj_c
    torch.cuda.empty_cache()
    return {"what": "oracle/torch_ref.py on cuda: the reference module's operators as torch 2.11 runs them (COO tensors, "
                    "coalesce + cuSPARSE SpMM per call, unfused elementwise kernels, autograd); same graph, same shapes",
            "ms_per_step": ms_step, "value": 1e3 / ms_step, "unit": "epochs/s",
            "spmm_F%d_ms" % hidden: ms_spmm, "spmm_F%d_ms_coalesced_input" % hidden: ms_spmm_c, "steps": steps, "warmup": warmup}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def kernel_name(csr, hidden: int, B=None) -> str:
    """The kernel the F = hidden product runs on this plan (asked of the library: tg_plan_spmm_launches)."""
    import torch
    probe = B if B is not None else torch.empty((2, hidden), dtype=torch.float32, device=csr.device)
    if csr.spmm_launches(probe, hidden) >= 2:
        return (f"roles2_kernel<GS={csr.hub_gs},NQ={csr.doc_nq}> + stream_finish_kernel (role-specialised column-chunk streaming SpMM: "
                f"TMA-staged {csr.chunk_rows}-node tiles, {csr.hub_groups} hub slot group(s), bulk-copy entry ring, FFMA2; F={hidden})")
    return f"spmm_kernel<4,32,{(hidden // 4 + 31) // 32},EpiStore> (gather SpMM, F={hidden})"


def measure_workload(name: str, args, world: int, rank: int, dev, hook, steps: int, warmup: int, extras: bool):
    """Builds one workload (single graph at world = 1, one shard per rank otherwise), times `steps` train steps after
    `warmup` with device-resident inputs and again end to end with host labels, and returns the measurements (every rank
    runs this; the timing is the max over ranks).  `extras`: also time the step with Adam and replayed from a CUDA graph."""
    import torch
    import torch.distributed as dist

    import topicgcn_b200 as tg
    from topicgcn_b200 import graphgen, ops

    hidden, n_class = graphgen.CONFIGS[name][2], graphgen.CONFIGS[name][3]
    g = None
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

        shard_desc = (f"documents row-sharded over {world} ranks, topic rows replicated; per step 3 NCCL all-reduces "
                      f"(K x H, K x C, packed [K x H | dW2 | db1 | db2 | loss]), two of them on a side stream")
    h2d = labels_host.numel() * 8 + index_host.numel() * 8
    params = list(model.parameters())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, w, use_hook=False):
        for _ in range(w):
            fn()
        barrier()
        ops.Stats.launches = 0
        hook.enabled = use_hook
        sampler = ClockSampler(dev.index or 0) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
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

    hook.records.clear()
    hook.dense.clear()
    # ---- device-resident timing (value) + live per-kernel events ------------------------------------------------------
    ms_total, launches, clocks = timed(step_device, steps, warmup, use_hook=True)
    ms_step = ms_total / steps
    out = {"name": name, "hidden": hidden, "classes": n_class, "ms_per_step": ms_step, "launches": launches, "clocks": clocks,
           "h2d": h2d, "n_docs_total": n_docs_total, "nnz_total": nnz_total, "shard_desc": shard_desc,
           "w1_gb": csr.n_cols * hidden * 4 / 1e9, "graph": g}
    # ---- roofline of the dominant kernel: the F = hidden SpMM (layer-1 forward and the dW1 backward) ----------------
    hbm_peak, peak_src = measured_peaks()
    kern = {}
    for (tag, f, _cid), (kcsr, ts) in hook.summary().items():
        if f is None or kcsr is None or kcsr.n_rows != csr.n_rows:
            continue  # (the sharded mode also runs the epilogue on the K replicated topic rows: not the SpMM)
        kern.setdefault((tag, f), []).append((kcsr, ts))
    detail = {}
    for (tag, f), lst in kern.items():
        ts = [t for _, tl in lst for t in tl]
        kcsr = lst[0][0]
        nbytes = spmm_bytes(kcsr.n_rows, kcsr.n_cols, kcsr.nnz, f)
        avg_ms = sum(ts) / len(ts)
        detail[f"{tag}_F{f}"] = {"launches": len(ts), "avg_ms": avg_ms, "algorithmic_GB": nbytes / 1e9,
                                 "GBps": nbytes / 1e6 / avg_ms, "frac_of_peak": nbytes / 1e6 / avg_ms / hbm_peak}
    # the skinny dense kernels of the step: bytes = the [n x h] operand once (dense_nn) / read + written once (hidden_bwd)
    for (tag, n_r, h_r, c_r), ts in hook.dense_summary().items():
        if n_r != csr.n_rows:
            continue  # (sharded mode: the product on the K replicated topic rows)
        nbytes = (1.0 if tag == "dense_nn" else 2.0) * n_r * h_r * 4.0 + n_r * c_r * 4.0 + h_r * c_r * 4.0
        avg_ms = sum(ts) / len(ts)
        detail[f"{tag}_{h_r}x{c_r}"] = {"launches": len(ts), "avg_ms": avg_ms, "algorithmic_GB": nbytes / 1e9,
                                         "GBps": nbytes / 1e6 / avg_ms, "frac_of_peak": nbytes / 1e6 / avg_ms / hbm_peak}
    out["kernels"] = detail
    roof = None
    dom = [(k, v) for k, v in detail.items() if k.endswith(f"_F{hidden}")]
    if dom:
        tot_ms = sum(v["avg_ms"] * v["launches"] for _, v in dom)
        tot_bytes = sum(v["algorithmic_GB"] * v["launches"] for _, v in dom)
        n_l = sum(v["launches"] for _, v in dom)
        achieved = tot_bytes * 1e3 / tot_ms
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if world == 1 and os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic = tj.get(name)  # dram bytes per launch from the committed ncu --set full capture of this kernel
            if not isinstance(traffic, (int, float)):
                traffic = None
            traffic_src = "static: profiles/roofline_traffic.json (dram__bytes of a committed ncu --set full capture, not measured in this run)"
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name(csr, hidden),
                "launches_timed": n_l, "avg_launch_ms": tot_ms / n_l, "algorithmic_bytes_per_launch": tot_bytes * 1e9 / n_l,
                "share_of_step": tot_ms / (ms_step * steps), "peak_source": peak_src,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "bytes_formula": "nnz*8 + (n_rows+1)*4 + n_cols*F*4 + n_rows*F*4 (SURVEY 8d)"}
    out["roofline"] = roof
    if extras:
        # ---- with Adam (reported, not the headline) -------------------------------------------------------------------
        opt = tg.optim.Adam(params, lr=0.02)  # one tg_adam_f32 pass per parameter (torch.optim.Adam semantics, trainer.py:307)

        def step_adam():
            step_device()
            opt.step()

        k_adam = max(3, steps // 2)
        ms_adam, _, _ = timed(step_adam, k_adam, 3)
        out["ms_adam"] = ms_adam / k_adam
        del opt
        # ---- the same step replayed from a CUDA graph (launch-bound small graphs; reported, not the headline) -------
        graph_info = None
        if world == 1:
            try:
                cap = tg.CapturedTrainStep(model, x, adj, g.labels, g.train_idx)
                ms_g, _, _ = timed(cap.step, steps, warmup)
                graph_info = {"ms_per_step": ms_g / steps, "value": 1e3 * steps / ms_g}
                del cap
                model._offset_dev = None
            except Exception as exc:  # pragma: no cover
                graph_info = {"error": str(exc)[:200]}
        out["cuda_graph_step"] = graph_info
    # ---- end to end through the public API with host inputs -----------------------------------------------------------
    ms_e2e, _, _ = timed(step_e2e, steps, warmup)
    out["ms_e2e"] = ms_e2e / steps
    # ---- N > 1: the replicated gradients must be bit-identical on every rank --------------------------------------------
    if world > 1:
        step_device()
        D = model.lg.n_docs_local
        rep = torch.cat([model.gc1.weight.grad[D:].reshape(-1), model.gc1.bias.grad, model.gc2.weight.grad.reshape(-1),
                         model.gc2.bias.grad]).contiguous().view(torch.int32)
        hi, lo = rep.clone(), rep.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        out["replicated_grads_bit_identical"] = bool(torch.equal(hi, lo))
    del model, params
    torch.cuda.empty_cache()
    return out


def sharded_vs_single_gpu_check(world: int, rank: int, dev):
    """Cross-rank correctness inside the benchmark run (real NCCL ranks): one train step, dropout off, on a small GLOBAL
    graph (20 K documents x 64 topics) cut into `world` document shards against the same step on one GPU (rank 0 runs the
    un-sharded module).  Returns the relative differences of the loss and of the replicated gradients."""
    import torch
    import torch.distributed as dist

    import topicgcn_b200 as tg
    from topicgcn_b200 import graphgen, ops, shard

    Dg, K, H, C = 20_000, 64, 64, 8
    gen = torch.Generator(device=dev).manual_seed(5)
    d, t, w = graphgen.doc_topic_edges(Dg, K, 2, 9, gen, dev)
    ti, tj, ts = graphgen.topic_topic_edges(K, gen, dev, dense=True)
    cg = torch.Generator(device="cpu").manual_seed(9)
    labels = torch.randint(0, C, (Dg,), generator=cg).to(dev)
    train_idx = torch.sort(torch.randperm(Dg, generator=cg)[: int(0.6 * Dg)]).values.to(dev)
    torch.manual_seed(3)
    full = tg.GCN(Dg + K, H, C, 0.0)
    params = {k: v.to(dev) for k, v in full.state_dict().items()}
    comm = shard.TorchDistComm()
    lo, hi = shard.shard_edges(Dg, world)[rank]
    mine = (d >= lo) & (d < hi)
    lg = shard.build_local_graph(d[mine] - lo, t[mine], w[mine], ti, tj, ts, hi - lo, K, comm)
    lg.n_train_global = int(train_idx.numel())
    model = shard.ShardedGCN(lg, H, C, 0.0, comm=comm).to(dev)
    with torch.no_grad():
        model.gc1.weight.copy_(torch.cat([params["gc1.weight"][lo:hi], params["gc1.weight"][Dg:]]))
        model.gc1.bias.copy_(params["gc1.bias"]); model.gc2.weight.copy_(params["gc2.weight"]); model.gc2.bias.copy_(params["gc2.bias"])
    model.train()
    tr_local = train_idx[(train_idx >= lo) & (train_idx < hi)] - lo
    loss = model.loss(row_label=ops.make_row_label(lg.n_local, labels[lo:hi], tr_local))
    loss.backward()
    res = None
    if rank == 0:
        u = torch.cat([d, ti + Dg]); v = torch.cat([t + Dg, tj + Dg]); ww = torch.cat([w, ts])
        r, c, vals = graphgen.normalize_undirected(u, v, ww, Dg + K)
        adj = torch.sparse_coo_tensor(torch.stack([r, c]), vals, (Dg + K, Dg + K), check_invariants=False)
        ref = full.to(dev)
        ref.train()
        ref_loss = ref.loss(tg.Featureless(Dg + K), adj, labels, train_idx)
        ref_loss.backward()

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

        res = {"graph": f"{Dg} documents x {K} topics over {world} NCCL ranks vs one GPU, dropout off",
               "loss_rel_err": abs(float(loss) - float(ref_loss)) / max(abs(float(ref_loss)), 1e-30),
               "grad_rel_err": {"gc1.weight[topic rows]": rel(model.gc1.weight.grad[hi - lo:], ref.gc1.weight.grad[Dg:]),
                                "gc1.weight[rank-0 document rows]": rel(model.gc1.weight.grad[:hi - lo], ref.gc1.weight.grad[lo:hi]),
                                "gc1.bias": rel(model.gc1.bias.grad, ref.gc1.bias.grad),
                                "gc2.weight": rel(model.gc2.weight.grad, ref.gc2.weight.grad),
                                "gc2.bias": rel(model.gc2.bias.grad, ref.gc2.bias.grad)}}
        res["ok"] = bool(res["loss_rel_err"] <= 1e-5 and max(res["grad_rel_err"].values()) <= 2e-5)
    dist.barrier()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    import topicgcn_b200 as tg
    from topicgcn_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line of the contract: NCCL prints its version banner on file descriptor 1 when the
        # communicator is created, so descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    tg._native.lib()
    hook = EventHook({"spmm", "gc1_fwd", "gc2_loss_fwd", "dense_nn", "hidden_bwd"})
    ops.set_kernel_hook(hook)

    name = WORKLOADS[args.workload]
    m = measure_workload(name, args, world, rank, dev, hook, args.steps, args.warmup, extras=True)
    shards = world  # one shard of the workload per rank
    value = shards * 1e3 / m["ms_per_step"]

    lib_base = None
    if world == 1 and rank == 0 and not args.no_library_baseline and m["graph"] is not None:
        try:
            lib_base = gpu_library_baseline(m["graph"], m["hidden"], m["classes"], dev)
            lib_base["ours_over_library"] = lib_base["ms_per_step"] / m["ms_per_step"]
        except Exception as exc:  # pragma: no cover
            lib_base = {"error": str(exc)[:300]}
    m["graph"] = None

    # ---- N > 1: the north-star configuration (BASELINE.json configs[3]) and the cross-rank checks ------------------------
    c4, consistency = None, None
    if world > 1:
        consistency = {"replicated_grads_bit_identical": m.get("replicated_grads_bit_identical")}
        try:
            chk = sharded_vs_single_gpu_check(world, rank, dev)
            if chk is not None:
                consistency["sharded_vs_single_gpu"] = chk
        except Exception as exc:  # pragma: no cover
            consistency["sharded_vs_single_gpu"] = {"error": str(exc)[:300]}
        if not args.no_c4 and args.workload != "c4":
            try:
                c = measure_workload(WORKLOADS["c4"], args, world, rank, dev, hook, args.c4_steps, 3, extras=False)
                c4 = {"workload": c["name"], "docs_per_gpu": c["n_docs_total"] // world, "docs_total": c["n_docs_total"],
                      "nnz_total": c["nnz_total"], "steps": args.c4_steps, "warmup": 3, "ms_per_step": c["ms_per_step"],
                      "epochs_per_sec_whole_graph": 1e3 / c["ms_per_step"],
                      "shard_epochs_per_sec_x_shards": world * 1e3 / c["ms_per_step"],
                      "e2e_ms_per_step": c["ms_e2e"], "gpu_launches": c["launches"], "roofline": c["roofline"],
                      "kernels": c["kernels"], "replicated_grads_bit_identical": c.get("replicated_grads_bit_identical"),
                      "scaling_note": "weak: 6.25 M documents per GPU at every N (50 M documents at N = 8); compare "
                                      "shard_epochs_per_sec_x_shards across N, or ms_per_step against the N = 1 run of --workload c4"}
            except Exception as exc:  # pragma: no cover
                c4 = {"error": str(exc)[:300]}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, desc, threads, ms = cpu_epoch_rate(args.workload, args.cpu_sample_docs, 3, 1)
            cpu = {"value": rate, "unit": "epochs/s", "cores": threads, "kind": "port", "sample": desc,
                   "ms_per_step_on_sample": ms, "value_from": "openmp_port"}
            try:  # the reference's own operators on the same slice (oracle/torch_ref.py): report the faster CPU leg
                ms_a, _, _ = aten_cpu_step_ms(args.workload, args.cpu_sample_docs, 3, 1)
                cpu["legs"] = {"openmp_port": {"ms_per_step_on_sample": ms}, "aten_operators": {"ms_per_step_on_sample": ms_a}}
                if ms_a < ms:
                    cpu.update(value=rate * ms / ms_a, ms_per_step_on_sample=ms_a, value_from="aten_operators")
            except Exception as exc:  # pragma: no cover
                cpu["legs"] = {"aten_operators": {"error": str(exc)[:200]}}
        line = {
            "metric": METRIC, "value": value, "unit": "epochs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "docs_total": m["n_docs_total"], "nnz_total": m["nnz_total"], "hidden": m["hidden"],
                       "classes": m["classes"], "featureless": True, "optimizer_in_step": False,
                       "l2_policy": "inputs larger than L2 (W1 alone is %.2f GB per GPU)" % m["w1_gb"],
                       "parallelism": m["shard_desc"],
                       "value_definition": "epochs/s of one shard x number of per-GPU shards (weak scaling: every GPU holds one "
                                           "graph of the workload's size; N = 1: plain epochs/s)"},
            "clocks": m["clocks"], "gpu_launches": m["launches"],
            "e2e": {"value": shards * 1e3 / m["ms_e2e"], "unit": "epochs/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4,
                    "ms_per_step": m["ms_e2e"]},
            "roofline": m["roofline"], "kernels": m["kernels"], "cuda_graph_step": m.get("cuda_graph_step"),
            "with_adam": {"ms_per_step": m["ms_adam"], "value": shards * 1e3 / m["ms_adam"]},
            "cpu_baseline": cpu, "gpu_library_baseline": lib_base, "c4": c4, "consistency": consistency,
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
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the torch.spmm-on-cuda (cuSPARSE) arm at N = 1")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the block on the 6.25 M x 1 024 per-GPU shard")
    ap.add_argument("--c4-steps", type=int, default=5)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
