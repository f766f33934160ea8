#!/usr/bin/env python
"""Measured TF32 tensor-core rate of this GPU (cuBLAS through torch, allow_tf32) and what it implies for a densified hub role.

The hub role of the wide SpMM computes, per (chunk, 128-column slice), P[256 slots x 128] += Theta^T[256 x 192] * B[192 x 128]
with ~8 stored entries per document.  A tensor-core version must treat Theta^T as DENSE and, to keep the 1e-5 parity
budget, run the 3xTF32 split (hi*hi + hi*lo + lo*hi): 3 * 2 * 256 * 128 * 192 flop per chunk-slice.  This script measures
the TF32 GEMM rate the library reaches on this box (an upper bound for any hand-written tcgen05 kind::tf32 kernel) and prints
the time such a chunk-slice would take on one SM at that rate next to the measured time of the CUDA-core hub role."""
import json
import sys

import torch


def rate(n=8192, reps=10):
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    tf = rate()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    flop = 3 * 2 * 256 * 128 * 192          # one (chunk, slice) of the C3 plan, 3xTF32, Theta^T densified
    us_tc = flop / (tf * 1e12 / sms) * 1e6
    out = {"tf32_gemm_tflops_measured": tf, "sms": sms, "densified_hub_chunk_slice_flop_3xtf32": flop,
           "us_per_chunk_slice_on_one_sm_at_that_rate": us_tc,
           "us_per_chunk_slice_cuda_core_hub_role_measured": 5.3,
           "note": "C3: 0.788 ms / (5210 chunks / 35 chunk lanes) = 5.3 us per chunk and slice on the FFMA2 hub role (profiles/r02_*); "
                   "the tensor-core figure excludes building the dense Theta^T tile and the hi/lo split of the B tile in shared memory"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
