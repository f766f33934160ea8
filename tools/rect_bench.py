#!/usr/bin/env python
"""Time the two products of a sparse feature matrix X [1M x 256, 50 entries per row] (X @ W and X^T @ dS) on the
rectangular sub-plans of the role kernels against the gather kernel (TG_ROLES2_RECT=0).  GPU box: python tools/rect_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import topicgcn_b200 as tg
dev = torch.device('cuda:0')
n, F, H, nnz_row = 1_000_256, 256, 256, 50
gen = torch.Generator(device=dev).manual_seed(1)
cols = torch.rand(n, F, device=dev, generator=gen).topk(nnz_row, dim=1).indices.sort(dim=1).values
vals = torch.rand(n, nnz_row, device=dev, generator=gen)
rows = torch.arange(n, device=dev).unsqueeze(1).expand(-1, nnz_row)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for rect in ("1", "0"):
    os.environ["TG_ROLES2_RECT"] = rect
    X = tg.DeviceCSR.from_coo(rows.reshape(-1), cols.reshape(-1), vals.reshape(-1), n, F)
    XT = X.transpose()
    W = torch.randn(F, H, device=dev); dS = torch.randn(n, H, device=dev)
    print("rect", rect, X.roles2_rect, XT.roles2_rect, "X@W %.3f ms" % t(lambda: tg.spmm(X, W)), "XT@dS %.3f ms" % t(lambda: tg.spmm(XT, dS)), flush=True)
