"""oracle/torch_ref.py — the reference's 2-layer GCN restated on the SAME ATen operators it runs on.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and by the baseline legs of bench.py,
never by the product package.

The reference's arithmetic for this path lives in PyTorch itself: `th.spmm(sparse_coo, dense)` twice per layer
(reference layer.py:102 feature transform, layer.py:106 aggregation), bias add (layer.py:109-110), `th.relu`
(layer.py:182), `th.dropout` (layer.py:185), `CrossEntropyLoss` on the training rows (trainer.py:308, 358-359) and
autograd for the backward.  The Python reference cannot travel to the GPU box, but these operators are
there (torch 2.11): this module strings the same calls together so that

  * `bench.py --impl reference` can time the ATen CPU path next to the OpenMP port (anchoring the port), and
  * `bench.py` can report the unmodified library path on the GPU (torch's COO -> coalesce -> cuSPARSE SpMM, unfused
    elementwise kernels): the bar SURVEY §2.1 names.

Checked against the real reference modules in this container by tests/test_oracle.py::test_torch_ref_matches_reference
(same seeds -> same initial weights, same logits / loss / gradients, bit for bit on CPU).
"""
from __future__ import annotations

import math

import torch as th


class GraphConvolutionRef(th.nn.Module):
    """out = adj @ (x @ W) + b with both products as sparse x dense `th.spmm` (layer.py:84-112)."""

    def __init__(self, n_in: int, n_out: int):
        super().__init__()
        self.weight = th.nn.Parameter(th.empty(n_in, n_out))
        self.bias = th.nn.Parameter(th.empty(n_out))
        bound = 1.0 / math.sqrt(n_out)  # layer.py:67-82: U(-1/sqrt(out), 1/sqrt(out)), weight first
        self.weight.data.uniform_(-bound, bound)
        self.bias.data.uniform_(-bound, bound)

    def forward(self, x, adj):
        support = th.spmm(x, self.weight) if x.is_sparse else th.mm(x, self.weight)
        return th.spmm(adj, support) + self.bias


class GCNRef(th.nn.Module):
    """gc1 -> relu -> dropout -> gc2 (layer.py:164-190); layer 2 sees a dense input, so its feature transform is the
    dense product the reference's `th.spmm(dense, dense)` call reduces to."""

    def __init__(self, nfeat: int, nhid: int, nclass: int, dropout: float):
        super().__init__()
        self.gc1 = GraphConvolutionRef(nfeat, nhid)
        self.gc2 = GraphConvolutionRef(nhid, nclass)
        self.dropout = dropout

    def forward(self, x, adj):
        h = th.relu(self.gc1(x, adj))
        h = th.dropout(h, self.dropout, train=self.training)
        return self.gc2(h, adj)


def sparse_identity(n: int, device) -> th.Tensor:
    """The featureless input the reference feeds layer 1 (X = I as a sparse COO tensor)."""
    ar = th.arange(n, device=device)
    return th.sparse_coo_tensor(th.stack([ar, ar]), th.ones(n, device=device), (n, n))


def train_step(model: GCNRef, x, adj, target: th.Tensor, index: th.Tensor) -> th.Tensor:
    """One iteration of the reference loop body without the optimizer (trainer.py:354-361): forward, loss on the
    training rows, backward."""
    for p in model.parameters():
        p.grad = None
    logits = model(x, adj)
    loss = th.nn.functional.cross_entropy(logits[index], target[index])
    loss.backward()
    return loss
