"""Drop-in `GraphConvolution` / `GCN` modules (reference layer.py:24-190) running on the topicgcn CUDA library.

Same constructor arguments, same parameter names / shapes / creation order / init distribution (so a seed gives the
same initial weights and `state_dict`s interchange with the reference modules), same `forward(x, adj)` signature:

    adj : torch.sparse COO  Â = D^-1/2 (A+I) D^-1/2 as produced by utils.preprocess_adj (reference utils.py:185-203),
          on the model's CUDA device.  Converted once to a device CSR and cached (csr.cached_csr).
    x   : layer-1 features.  torch.sparse COO [N x nfeat] (the reference's mode, trainer.py:238,344), a dense tensor,
          or the featureless identity (a sparse identity matrix, `Featureless(n)`, or None): then X·W1 is W1 itself
          and dW1 is the aggregated gradient — no product is executed.

`GCN.forward` returns logits for all N rows like the reference (layer.py:190).  `GCN.loss` is the fused train-step
entry (forward + masked cross-entropy of trainer.py:357-359 in one autograd node).
"""
from __future__ import annotations

import functools

import math
from typing import Optional

import torch
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from . import _native as N
from . import ops
from .csr import DeviceCSR, cached_csr


class Featureless:
    """Marker for X = I (the TextGCN featureless mode, reference layer.py:134 docstring) without materialising
    an N x N sparse identity."""

    def __init__(self, n: int):
        self.n = int(n)

    @property
    def shape(self):
        return (self.n, self.n)

    def to(self, *_a, **_k):
        return self


_identity_memo: dict = {}


def _is_identity(x) -> bool:
    """True when x is the featureless input.  A sparse COO tensor is inspected once (one host sync) and memoised."""
    if x is None or isinstance(x, Featureless):
        return True
    if not isinstance(x, torch.Tensor) or x.layout != torch.sparse_coo:
        return False
    if x.shape[0] != x.shape[1] or x._nnz() != x.shape[0]:
        return False
    key = (id(x), x._values().data_ptr(), x._values()._version, x._indices()._version, x._nnz())
    hit = _identity_memo.get(key)
    if hit is None:
        idx, val = x._indices(), x._values()
        ar = torch.arange(x.shape[0], device=idx.device)
        hit = bool(((idx[0] == ar) & (idx[1] == ar)).all().item() and (val == 1).all().item())
        if len(_identity_memo) > 64:
            _identity_memo.clear()
        _identity_memo[key] = hit
    return hit


def _as_csr(adj) -> DeviceCSR:
    if isinstance(adj, DeviceCSR):
        return adj
    if isinstance(adj, torch.Tensor) and adj.layout in (torch.sparse_coo, torch.sparse_csr):
        if not adj.is_cuda:
            raise N.TopicGCNError("adj must be on a CUDA device; topicgcn_b200 has no CPU fallback")
        return cached_csr(adj)
    raise N.TopicGCNError("adj must be a torch sparse tensor (COO/CSR) or a DeviceCSR")


def _metrics_from_counts(tp, fp, fn, n_rows: int) -> dict:
    """utils.accuracy / utils.macro_f1 (reference utils.py:71-85, 107): per-class ratios with 0/0 -> 0, macro-averaged,
    F1 of the two averages."""
    import numpy as np
    tp, fp, fn = (np.asarray(a, dtype=np.float64) for a in (tp, fp, fn))
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = tp / (tp + fp)
        rec = tp / (tp + fn)
    prec[np.isnan(prec)] = 0
    rec[np.isnan(rec)] = 0
    P, R = float(prec.mean()), float(rec.mean())
    with np.errstate(divide="ignore", invalid="ignore"):
        f1 = float(np.float64(2 * P * R) / np.float64(P + R))
    return {"acc": float(tp.sum()) / max(int(n_rows), 1), "macro_f1": f1, "precision": P, "recall": R}


def _support(x, weight: torch.Tensor) -> torch.Tensor:
    """support = X @ W  (reference layer.py:102) for the three kinds of X."""
    if _is_identity(x):
        if x is not None and x.shape[1] != weight.shape[0]:
            raise N.TopicGCNError("featureless input needs nfeat == number of nodes")
        return weight
    if isinstance(x, DeviceCSR):
        return ops.SpMMFunction.apply(weight, None, x)
    if x.layout in (torch.sparse_coo, torch.sparse_csr):
        return ops.SpMMFunction.apply(weight, None, cached_csr(x))
    # dense features (not the reference's mode: its features are sparse): tg_gemm_f32, dW1 = X^T dS1 in the backward
    if not x.is_cuda:
        raise N.TopicGCNError("x must be on a CUDA device; topicgcn_b200 has no CPU fallback")
    return ops.DenseFeatureTransform.apply(x, weight)


class GraphConvolution(Module):
    """One graph-convolution layer  out = Â (X W) + b   (reference layer.py:24-123)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        # reference layer.py:67-82: U(-1/sqrt(out), 1/sqrt(out)) for the weight, then the bias (same RNG order)
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, infeatn, adj):
        support = _support(infeatn, self.weight)
        return ops.SpMMFunction.apply(support, self.bias, _as_csr(adj))

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


class GCN(Module):
    """Two-layer GCN: gc1 -> relu -> dropout -> gc2 (reference layer.py:126-190)."""

    def __init__(self, nfeat: int, nhid: int, nclass: int, dropout: float):
        super().__init__()
        self.gc1 = GraphConvolution(nfeat, nhid)
        self.gc2 = GraphConvolution(nhid, nclass)
        self.dropout = dropout
        # dropout RNG: counter-based Philox keyed on (seed, call counter); the seed is drawn from torch's global
        # generator on first use so `th.manual_seed(seed)` (trainer.py:294-296) makes runs reproducible
        self._dropout_seed: Optional[int] = None
        self._dropout_calls = 0
        self._next_keep_mask: Optional[torch.Tensor] = None
        self._offset_dev: Optional[torch.Tensor] = None  # device-side call counter (CUDA-graph replays, graph.py)

    # ---- dropout control -------------------------------------------------------------------------------------
    def set_dropout_seed(self, seed: int) -> None:
        self._dropout_seed, self._dropout_calls = int(seed), 0

    def set_next_dropout_mask(self, keep_mask: Optional[torch.Tensor]) -> None:
        """Parity mode: use this explicit uint8 keep mask [N x nhid] for the next training forward (e.g. the mask
        torch's `bernoulli_(1-p)` drew for the reference run), instead of the Philox stream."""
        self._next_keep_mask = keep_mask

    def _dropout_state(self):
        mask = self._next_keep_mask
        self._next_keep_mask = None
        if not self.training or self.dropout <= 0.0:
            return None, 0, 0
        if self._dropout_seed is None:
            self._dropout_seed = int(torch.randint(0, 2**62, (1,)).item())
        if self._offset_dev is None:
            self._dropout_calls += 1  # (with a device counter the host-side offset stays fixed: CUDA-graph replays)
        return mask, self._dropout_seed, self._dropout_calls

    # ---- reference API --------------------------------------------------------------------------------------
    def forward(self, x, adj):
        csr = _as_csr(adj)
        S1 = _support(x, self.gc1.weight)
        mask, seed, off = self._dropout_state()
        return ops.GCNCoreFunction.apply(S1, self.gc1.bias, self.gc2.weight, self.gc2.bias, csr, float(self.dropout),
                                         bool(self.training), mask, seed, off, self._offset_dev)

    # ---- fused train-step API -----------------------------------------------------------------------------------
    def loss(self, x, adj, target: torch.Tensor, index: torch.Tensor, return_logits: bool = False,
             row_label: Optional[torch.Tensor] = None):
        """mean cross-entropy of `forward(x, adj)[index]` against `target[index]` — what the reference computes at
        trainer.py:357-359 — with bias + log-softmax + NLL and its gradient fused into the layer-2 SpMM epilogue."""
        csr = _as_csr(adj)
        S1 = _support(x, self.gc1.weight)
        mask, seed, off = self._dropout_state()
        if row_label is None:
            if target.is_cuda and index.is_cuda:
                row_label = ops.make_row_label(csr.n_rows, target, index)
            else:  # host labels / indices (trainer.py keeps them on the host until convert_tensor): copy on a side stream
                # (a callable: the fused step launches layer 1 first and only then queues the copies on the side stream)
                row_label = functools.partial(ops.make_row_label_async, csr.n_rows, target, index, self.gc1.weight.device)
        inv = 1.0 / max(int(index.numel()), 1)
        loss, logits = ops.GCNLossFunction.apply(S1, self.gc1.bias, self.gc2.weight, self.gc2.bias, csr,
                                                 float(self.dropout), bool(self.training), mask, seed, off, row_label,
                                                 inv, bool(return_logits), self._offset_dev)
        return (loss, logits) if return_logits else loss

    @torch.no_grad()
    def evaluate(self, x, adj, target: torch.Tensor, index: torch.Tensor, prefix: str = "val") -> dict:
        """The reference's validation / test pass (`TopicGCNTrainer.val`, trainer.py:378-398): eval-mode forward, mean
        cross-entropy, accuracy and macro F1 / precision / recall on `index` — with the per-class counts taken by one
        kernel and ONE device->host copy instead of the reference's 3 * nclass + 2 `.item()` syncs (utils.py:25-109).
        The module's train / eval flag is left as it was.  Returns the reference's dict."""
        was_training = self.training
        self.eval()
        try:
            csr = _as_csr(adj)
            S1 = _support(x, self.gc1.weight)
            if target.is_cuda and index.is_cuda:
                row_label = ops.make_row_label(csr.n_rows, target, index)
            else:
                row_label = ops.make_row_label_async(csr.n_rows, target, index, self.gc1.weight.device)
            n_index = max(int(index.numel()), 1)
            H1 = ops.gc1_forward(csr, S1.detach(), self.gc1.bias.detach() if self.gc1.bias is not None else None,
                                 float(self.dropout), False)
            S2 = ops.dense_nn(H1, self.gc2.weight.detach())
            loss, logits, _ = ops.gc2_loss_forward(csr, S2, self.gc2.bias.detach() if self.gc2.bias is not None else None,
                                                   row_label, 1.0 / n_index, want_logits=True, want_grad=False)
            counts = ops.class_counts(logits, row_label)
            packed = torch.cat([loss.reshape(1).to(torch.float64), counts.reshape(-1).to(torch.float64)]).cpu()  # the one sync
        finally:
            self.train(was_training)
        loss_v = float(packed[0])
        c = counts.shape[1]
        tp, fp, fn = (packed[1 + i * c:1 + (i + 1) * c].numpy() for i in range(3))
        return {f"{prefix}_loss": loss_v, **_metrics_from_counts(tp, fp, fn, n_index)}
