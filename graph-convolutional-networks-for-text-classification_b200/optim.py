"""Adam for the drop-in modules: `torch.optim.Adam` semantics (reference trainer.py:307, `th.optim.Adam(model.parameters(),
lr=0.02)`, stepped at trainer.py:362) with the update of every parameter done by ONE pass of `tg_adam_f32`.

With featureless input (X = I) the weight of layer 1 is [N x hidden] — 1 GB at 1 M nodes — and the optimizer step is a
pure streaming problem: 16 bytes read and 12 written per element.  State keys (`step`, `exp_avg`, `exp_avg_sq`) match
torch's, so `state_dict()`s interchange with `torch.optim.Adam`.  No amsgrad / maximize / capturable variants: the
reference uses none of them.  CUDA fp32 contiguous parameters only — anything else raises (no fallback).
"""
from __future__ import annotations

import torch

from . import _native as N


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = N.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise N.TopicGCNError("topicgcn_b200.optim.Adam needs contiguous CUDA fp32 parameters")
                if g.is_sparse or g.dtype != torch.float32:
                    raise N.TopicGCNError("topicgcn_b200.optim.Adam needs dense fp32 gradients")
                g = g.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)  # host scalar, like torch's non-capturable Adam
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                with torch.cuda.device(p.device):
                    N.check(lib.tg_adam_f32(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                            p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                            float(group["weight_decay"]), int(st["step"].item()),
                                            torch.cuda.current_stream(p.device).cuda_stream), "tg_adam_f32")
        return loss
