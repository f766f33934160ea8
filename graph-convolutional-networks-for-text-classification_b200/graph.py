"""CUDA-graph capture of the train step (forward + masked cross-entropy + backward).

On the small graphs (R8 shape: 7.7 K nodes, 20NG shape: 19 K nodes) the eleven kernels of a step take ~0.2 ms in
total while Python + launch overhead takes ~0.35 ms: the step is launch-bound.  `CapturedTrainStep` records the whole
step once into a `torch.cuda.CUDAGraph` and replays it per epoch (the reference's loop body, trainer.py:354-361).
Gradients land in the parameters' static `.grad` tensors (the kernels' own output buffers, rewritten by every replay), so
`optimizer.step()` after `step()` works unchanged; do not call `zero_grad(set_to_none=True)` between steps.
The dropout mask stays fresh across replays: the Philox call counter lives on the device (`offset_dev`,
include/topicgcn.h) and is bumped inside the graph.
"""
from __future__ import annotations

import torch

from . import ops
from .layer import GCN, _as_csr


class CapturedTrainStep:
    def __init__(self, model: GCN, x, adj, target: torch.Tensor, index: torch.Tensor, warmup: int = 3):
        self.model = model
        dev = next(model.parameters()).device
        csr = _as_csr(adj)
        csr.transpose()  # plan-time work (symmetry check) must not happen inside the capture
        self.row_label = ops.make_row_label(csr.n_rows, target, index)
        self.index = index
        model.train()
        if model._dropout_seed is None:
            model.set_dropout_seed(int(torch.randint(0, 2**62, (1,)).item()))
        model._offset_dev = torch.zeros((), dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # allocates workspaces and the static .grad tensors
                for p in model.parameters():
                    if p.grad is not None:
                        p.grad.zero_()
                model.loss(x, adj, target, index, row_label=self.row_label).backward()
                model._offset_dev.add_(1)
        torch.cuda.current_stream(dev).wait_stream(side)
        model._dropout_calls = 0  # from here on only the device counter advances
        self.graph = torch.cuda.CUDAGraph()
        # Inside the capture the parameters carry NO gradient: autograd then installs the kernels' output buffers as the
        # `.grad` tensors (no zero-fill, no accumulation pass — 3 GB of traffic for the 1 GB featureless weight), and
        # because those buffers live in the graph's private pool every replay rewrites the same tensors in place.
        for p in model.parameters():
            p.grad = None
        with torch.cuda.graph(self.graph):
            self.loss = model.loss(x, adj, target, index, row_label=self.row_label)
            self.loss.backward()
            model._offset_dev.add_(1)
        self.grads = [p.grad for p in model.parameters()]  # static: do not zero_grad(set_to_none=True) between steps

    def step(self) -> torch.Tensor:
        """One train step: replays the captured graph; returns the (static) loss tensor."""
        self.graph.replay()
        return self.loss
