"""Operator layer: thin Python wrappers over the C-ABI and the autograd Functions built from them.

Low-level wrappers (`spmm`, `gc1_forward`, `gc2_loss_forward`, `dense_nn`, `hidden_backward`, ...) map 1:1 onto
include/topicgcn.h.  The autograd Functions mirror what the reference gets from `th.spmm` + autograd:

  SpMMFunction        Y = A @ B (+bias)                 reference layer.py:106,110  (bwd: A^T @ dY)
  GCNCoreFunction     logits = A @ (dropout(relu(A @ S1 + b1)) @ W2) + b2      layer.py:181-188
  GCNLossFunction     the same + mean masked cross-entropy                      trainer.py:357-361
  MaskedCrossEntropy  loss on existing logits                                   trainer.py:358-359
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as N
from .csr import DeviceCSR


# ---------------------------------------------------------------------------------------------------------------
# instrumentation: how many of OUR kernels were launched, and an optional per-kernel CUDA-event hook (bench.py)
# ---------------------------------------------------------------------------------------------------------------
class Stats:
    launches = 0  # kernels of libtopicgcn.so launched through this module since the last reset


_kernel_hook = None


def set_kernel_hook(hook) -> None:
    """hook.start(tag, info) -> token / hook.stop(token) are called around every C-ABI compute call (bench.py records
    CUDA events on the launching stream with it).  None disables."""
    global _kernel_hook
    _kernel_hook = hook


class _call:
    """Context around one C-ABI call: counts its kernel launches and feeds the optional event hook."""

    def __init__(self, tag: str, launches: int, **info):
        self.tag, self.launches, self.info, self.tok = tag, launches, info, None

    def __enter__(self):
        Stats.launches += self.launches
        if _kernel_hook is not None:
            self.tok = _kernel_hook.start(self.tag, self.info)
        return self

    def __exit__(self, *exc):
        if _kernel_hook is not None and self.tok is not None:
            _kernel_hook.stop(self.tok)
        return False


def _spmm_launches(csr, B: torch.Tensor, n_feat: int, philox: bool = False, out: Optional[torch.Tensor] = None,
                   loss: bool = False) -> int:
    """Kernels one SpMM-type call launches: the product itself, the finishing kernel of the streaming paths (hub rows =
    fixed-order sum of the per-CTA partials + epilogue) and, for Philox dropout on the wide streaming path, the kernel
    that draws the bit-packed keep mask.  Asked of the library (tg_plan_spmm_launches), not mirrored here."""
    ok4 = out is None or (out.data_ptr() % 16 == 0 and (out.shape[0] <= 1 or out.stride(0) % 4 == 0))
    return csr.spmm_launches(B, n_feat, philox, ok4, loss=loss)


# ---------------------------------------------------------------------------------------------------------------
# tensor plumbing
# ---------------------------------------------------------------------------------------------------------------
def _dense2d(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise N.TopicGCNError(f"{what} must be a CUDA tensor (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise N.TopicGCNError(f"{what} must be float32 (got {t.dtype})")
    if t.dim() != 2:
        raise N.TopicGCNError(f"{what} must be 2-D")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _vec(t: Optional[torch.Tensor], n: int, what: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32 or t.numel() != n:
        raise N.TopicGCNError(f"{what} must be a CUDA float32 vector of {n} elements")
    return t.contiguous()


def _ld(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def _stream() -> int:
    return N.current_stream_ptr()


# ---------------------------------------------------------------------------------------------------------------
# 1:1 wrappers
# ---------------------------------------------------------------------------------------------------------------
def spmm(csr: DeviceCSR, B: torch.Tensor, bias: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, out_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = (A @ B (+ bias)) * out_scale: tg_spmm_f32.  out_scale is an optional 0-dim CUDA tensor (no host sync)."""
    B = _dense2d(B, "B")
    if B.shape[0] != csr.n_cols:
        raise N.TopicGCNError(f"shape mismatch: A is {csr.n_rows}x{csr.n_cols}, B has {B.shape[0]} rows")
    F = int(B.shape[1])
    bias = _vec(bias, F, "bias")
    if out is None:
        out = torch.empty((csr.n_rows, F), dtype=torch.float32, device=B.device)
    ws, ws_bytes = csr.workspace(F)
    with torch.cuda.device(B.device), _call("spmm", _spmm_launches(csr, B, F, out=out), n_feat=F, csr=csr):
        N.check(N.lib().tg_spmm_f32(csr.plan, N.ptr(csr.rowptr), N.ptr(csr.colidx), N.ptr(csr.vals), N.ptr(B), _ld(B),
                                    N.ptr(out), _ld(out), F, N.ptr(bias), N.ptr(out_scale), ws, ws_bytes, _stream()),
                "tg_spmm_f32")
    return out


def gc1_forward(csr: DeviceCSR, S: torch.Tensor, bias: Optional[torch.Tensor], p: float, training: bool,
                keep_mask: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0,
                out: Optional[torch.Tensor] = None, raw_row_begin: int = -1,
                offset_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """H1 = dropout(relu(A @ S + b1)): tg_gc1_fwd_f32.  Rows >= raw_row_begin (if >= 0) get the plain sums.
    offset_dev: optional int64 CUDA scalar added to `offset` on the device (CUDA-graph replays)."""
    S = _dense2d(S, "S")
    if S.shape[0] != csr.n_cols:
        raise N.TopicGCNError(f"shape mismatch: A is {csr.n_rows}x{csr.n_cols}, S has {S.shape[0]} rows")
    F = int(S.shape[1])
    bias = _vec(bias, F, "bias")
    if keep_mask is not None:
        if not keep_mask.is_cuda or keep_mask.dtype != torch.uint8 or tuple(keep_mask.shape) != (csr.n_rows, F):
            raise N.TopicGCNError("keep_mask must be a CUDA uint8 tensor of shape [n_rows, F]")
        keep_mask = keep_mask.contiguous()
    if out is None:
        out = torch.empty((csr.n_rows, F), dtype=torch.float32, device=S.device)
    ws, ws_bytes = csr.workspace(F)
    with torch.cuda.device(S.device), _call("gc1_fwd", _spmm_launches(csr, S, F, philox=bool(training) and keep_mask is None and p > 0.0, out=out), n_feat=F, csr=csr):
        N.check(N.lib().tg_gc1_fwd_f32(csr.plan, N.ptr(csr.rowptr), N.ptr(csr.colidx), N.ptr(csr.vals), N.ptr(S), _ld(S),
                                       N.ptr(bias), N.ptr(out), _ld(out), F, float(p), int(bool(training)),
                                       N.ptr(keep_mask), int(seed) & (2**64 - 1), int(offset) & (2**64 - 1),
                                       N.ptr(offset_dev), int(raw_row_begin), ws, ws_bytes, _stream()), "tg_gc1_fwd_f32")
    return out


def dropout_keep_mask(n_rows: int, n_feat: int, p: float, seed: int, offset: int, device) -> torch.Tensor:
    """The Philox keep mask tg_gc1_fwd_f32 uses for (seed, offset): tg_dropout_keep_mask."""
    out = torch.empty((n_rows, n_feat), dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device), _call("keep_mask", 1):
        N.check(N.lib().tg_dropout_keep_mask(N.ptr(out), n_rows, n_feat, float(p), int(seed) & (2**64 - 1),
                                             int(offset) & (2**64 - 1), _stream()), "tg_dropout_keep_mask")
    return out


def reduce_sum(x: torch.Tensor) -> torch.Tensor:
    """Deterministic sum of a float32 vector: tg_reduce_sum_f32."""
    x = x.contiguous().view(-1)
    n = int(x.numel())
    scratch = torch.empty(int(N.lib().tg_reduce_scratch_floats(n)), dtype=torch.float32, device=x.device)
    out = torch.empty((), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _call("reduce_sum", 2):
        N.check(N.lib().tg_reduce_sum_f32(N.ptr(x), n, N.ptr(scratch), N.ptr(out), _stream()), "tg_reduce_sum_f32")
    return out


def make_row_label(n_rows: int, target: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """row_label[i] = target[i] for i in `index`, -1 elsewhere (int32) — the masked-loss operand that replaces the
    reference's `logits[train_lst]` / `target[train_lst]` gathers (trainer.py:358-359).  Indices must be unique."""
    row_label = torch.full((n_rows,), -1, dtype=torch.int32, device=index.device)
    row_label[index] = target[index].to(torch.int32)
    return row_label


_side_streams: dict = {}
_ready_events: dict = {}


def make_row_label_async(n_rows: int, target: torch.Tensor, index: torch.Tensor, device) -> torch.Tensor:
    """make_row_label for labels / indices that live in (pinned) HOST memory: the host->device copies and the scatter run
    on a side stream, so they overlap the layer-1 kernels; the consumer (gc2_loss_forward / masked_ce) makes the compute
    stream wait on the recorded event just before it launches.  The host tensors must stay unchanged until then."""
    device = torch.device(device)
    side = _side_streams.get(device)
    if side is None:
        side = _side_streams[device] = torch.cuda.Stream(device)
    main = torch.cuda.current_stream(device)
    with torch.cuda.stream(side):
        t = target.to(device, non_blocking=True)
        i = index.to(device, non_blocking=True)
        row_label = make_row_label(n_rows, t, i)
        ev = torch.cuda.Event()
        ev.record(side)
    row_label.record_stream(main)
    _ready_events[row_label.data_ptr()] = ev
    return row_label


def _wait_ready(row_label: torch.Tensor) -> None:
    ev = _ready_events.pop(row_label.data_ptr(), None)
    if ev is not None:
        torch.cuda.current_stream(row_label.device).wait_event(ev)


def gc2_loss_forward(csr: DeviceCSR, S2: torch.Tensor, bias: Optional[torch.Tensor], row_label: torch.Tensor,
                     inv_count: float, want_logits: bool = True, want_grad: bool = True):
    """(loss, logits, dZ2) with Z2 = A @ S2 + b2 and loss = mean CE over labelled rows: tg_gc2_loss_fwd_f32."""
    S2 = _dense2d(S2, "S2")
    Cc = int(S2.shape[1])
    bias = _vec(bias, Cc, "bias")
    if row_label.dtype != torch.int32 or row_label.numel() != csr.n_rows or not row_label.is_cuda:
        raise N.TopicGCNError("row_label must be a CUDA int32 vector with one entry per row")
    dev = S2.device
    _wait_ready(row_label)
    logits = torch.empty((csr.n_rows, Cc), dtype=torch.float32, device=dev) if want_logits else None
    dZ2 = torch.empty((csr.n_rows, Cc), dtype=torch.float32, device=dev) if want_grad else None
    row_loss = torch.empty(csr.n_rows, dtype=torch.float32, device=dev)
    ws, ws_bytes = csr.workspace(Cc)
    with torch.cuda.device(dev), _call("gc2_loss_fwd", _spmm_launches(csr, S2, Cc, loss=True), n_feat=Cc, csr=csr):
        N.check(N.lib().tg_gc2_loss_fwd_f32(csr.plan, N.ptr(csr.rowptr), N.ptr(csr.colidx), N.ptr(csr.vals), N.ptr(S2),
                                            _ld(S2), N.ptr(bias), N.ptr(row_label), float(inv_count), N.ptr(logits),
                                            Cc, N.ptr(dZ2), Cc, N.ptr(row_loss), Cc, ws, ws_bytes, _stream()),
                "tg_gc2_loss_fwd_f32")
    return reduce_sum(row_loss), logits, dZ2


def masked_ce(logits: torch.Tensor, row_label: torch.Tensor, inv_count: float, want_grad: bool = True):
    """(loss, dZ) on existing logits: tg_masked_ce_f32."""
    logits = _dense2d(logits, "logits")
    n, Cc = int(logits.shape[0]), int(logits.shape[1])
    _wait_ready(row_label)
    dZ = torch.empty((n, Cc), dtype=torch.float32, device=logits.device) if want_grad else None
    row_loss = torch.empty(n, dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device), _call("masked_ce", 1):
        N.check(N.lib().tg_masked_ce_f32(N.ptr(logits), _ld(logits), N.ptr(row_label), float(inv_count), N.ptr(dZ), Cc,
                                         N.ptr(row_loss), n, Cc, _stream()), "tg_masked_ce_f32")
    return reduce_sum(row_loss), dZ


def class_counts(logits: torch.Tensor, row_label: torch.Tensor) -> torch.Tensor:
    """int32 [3 x C] true positives / false positives / false negatives of argmax(logits) on the rows with
    row_label >= 0: tg_class_counts_i32 (the sums behind utils.accuracy / utils.macro_f1, reference utils.py:25-109)."""
    logits = _dense2d(logits, "logits")
    n, Cc = int(logits.shape[0]), int(logits.shape[1])
    if row_label.dtype != torch.int32 or row_label.numel() != n or not row_label.is_cuda:
        raise N.TopicGCNError("row_label must be a CUDA int32 vector with one entry per row")
    _wait_ready(row_label)
    counts = torch.empty((3, Cc), dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device), _call("class_counts", 1):
        N.check(N.lib().tg_class_counts_i32(N.ptr(logits), _ld(logits), N.ptr(row_label), n, Cc, N.ptr(counts), _stream()),
                "tg_class_counts_i32")
    return counts


def dense_nn(A: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A[n x h] @ W[h x c] for skinny c: tg_dense_nn_f32.  `out`: optional [n x c] destination (contiguous rows)."""
    A = _dense2d(A, "A")
    W = _dense2d(W, "W")
    n, h, c = int(A.shape[0]), int(A.shape[1]), int(W.shape[1])
    if W.shape[0] != h:
        raise N.TopicGCNError("inner dimensions differ")
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=A.device)
    elif tuple(out.shape) != (n, c) or out.dtype != torch.float32 or not out.is_contiguous():
        raise N.TopicGCNError("out must be a contiguous float32 [n x c] tensor")
    with torch.cuda.device(A.device), _call("dense_nn", (c + 31) // 32, n=n, h=h, c=c):
        N.check(N.lib().tg_dense_nn_f32(N.ptr(A), _ld(A), N.ptr(W), _ld(W), N.ptr(out), c, n, h, c, _stream()),
                "tg_dense_nn_f32")
    return out


def colsum(X: torch.Tensor) -> torch.Tensor:
    """Deterministic column sums: tg_colsum_f32."""
    X = _dense2d(X, "X")
    n, c = int(X.shape[0]), int(X.shape[1])
    scratch = torch.empty(int(N.lib().tg_colsum_scratch_floats(n, c)), dtype=torch.float32, device=X.device)
    out = torch.empty(c, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device), _call("colsum", 2):
        N.check(N.lib().tg_colsum_f32(N.ptr(X), _ld(X), n, c, N.ptr(scratch), N.ptr(out), _stream()), "tg_colsum_f32")
    return out


def relu_dropout_backward(H: torch.Tensor, dH: torch.Tensor, scale: float) -> torch.Tensor:
    """dZ = dH * [H > 0] * scale: tg_relu_dropout_bwd_f32."""
    H = _dense2d(H, "H")
    dH = _dense2d(dH, "dH")
    n, f = int(H.shape[0]), int(H.shape[1])
    dZ = torch.empty((n, f), dtype=torch.float32, device=H.device)
    with torch.cuda.device(H.device), _call("relu_dropout_bwd", 1):
        N.check(N.lib().tg_relu_dropout_bwd_f32(N.ptr(H), _ld(H), N.ptr(dH), _ld(dH), float(scale), N.ptr(dZ), f, n, f,
                                                _stream()), "tg_relu_dropout_bwd_f32")
    return dZ


def hidden_backward(H1: torch.Tensor, dS2: torch.Tensor, W2: torch.Tensor, scale: float,
                    out_dZ1: Optional[torch.Tensor] = None, n_count: Optional[int] = None):
    """(dZ1, dW2, db1) — fused backward of S2 = H1 @ W2, dropout, relu, +b1: tg_hidden_bwd_rows_f32.  Any class count and
    hidden width: up to 32 classes (and 1024 hidden units) in one pass over H1, more in blocks of 32 classes / 256 units.
    n_count: rows from n_count on get their dZ1 but do not count into dW2 / db1 (default: every row counts)."""
    H1 = _dense2d(H1, "H1")
    dS2 = _dense2d(dS2, "dS2")
    W2 = _dense2d(W2, "W2")
    n, h, c = int(H1.shape[0]), int(H1.shape[1]), int(W2.shape[1])
    dev = H1.device
    dZ1 = out_dZ1 if out_dZ1 is not None else torch.empty((n, h), dtype=torch.float32, device=dev)
    dW2 = torch.empty((h, c), dtype=torch.float32, device=dev)
    db1 = torch.empty(h, dtype=torch.float32, device=dev)
    scratch = torch.empty(int(N.lib().tg_hidden_bwd_scratch_floats(n, h, c)), dtype=torch.float32, device=dev)
    blocks = 1 if (c <= 32 and h <= 1024) else ((c + 31) // 32) * ((h + 255) // 256)
    with torch.cuda.device(dev), _call("hidden_bwd", 2 * blocks, n=n, h=h, c=c):
        N.check(N.lib().tg_hidden_bwd_rows_f32(N.ptr(H1), _ld(H1), N.ptr(dS2), _ld(dS2), N.ptr(W2), _ld(W2), float(scale),
                                               N.ptr(dZ1), _ld(dZ1), N.ptr(dW2), N.ptr(db1), N.ptr(scratch), n, h, c,
                                               n if n_count is None else int(n_count), _stream()),
                "tg_hidden_bwd_rows_f32")
    return dZ1, dW2, db1


def gemm(A: torch.Tensor, B: torch.Tensor, trans_a: bool = False) -> torch.Tensor:
    """A @ B (trans_a = False) or A^T @ B (trans_a = True) for dense fp32 operands: tg_gemm_f32 — the products of a DENSE
    layer-1 feature matrix (reference layer.py:102 when `infeatn` is dense, and its autograd transpose product)."""
    A = _dense2d(A, "A")
    B = _dense2d(B, "B")
    k = int(B.shape[0])
    m = int(A.shape[1]) if trans_a else int(A.shape[0])
    if (int(A.shape[0]) if trans_a else int(A.shape[1])) != k:
        raise N.TopicGCNError("inner dimensions differ")
    n = int(B.shape[1])
    out = torch.empty((m, n), dtype=torch.float32, device=A.device)
    ns = int(N.lib().tg_gemm_scratch_floats(int(trans_a), m, n, k))
    scratch = torch.empty(ns, dtype=torch.float32, device=A.device) if ns else None
    with torch.cuda.device(A.device), _call("gemm", 2 if ns else 1, m=m, n=n, k=k):
        N.check(N.lib().tg_gemm_f32(int(trans_a), N.ptr(A), _ld(A), N.ptr(B), _ld(B), N.ptr(out), n, m, n, k, N.ptr(scratch),
                                    _stream()), "tg_gemm_f32")
    return out


# ---------------------------------------------------------------------------------------------------------------
# autograd
# ---------------------------------------------------------------------------------------------------------------
class SpMMFunction(torch.autograd.Function):
    """Y = A @ B (+ bias).  A is constant (the reference never differentiates adj or X)."""

    @staticmethod
    def forward(ctx, B, bias, csr: DeviceCSR):
        ctx.csr, ctx.has_bias = csr, bias is not None
        return spmm(csr, B, bias)

    @staticmethod
    def backward(ctx, dY):
        dY = dY.contiguous()
        dB = spmm(ctx.csr.transpose(), dY) if ctx.needs_input_grad[0] else None
        db = colsum(dY) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        return dB, db, None


class DenseFeatureTransform(torch.autograd.Function):
    """support = X @ W for a dense, constant feature matrix X (reference layer.py:102 with dense `infeatn`); only W gets a
    gradient (dW = X^T @ dS), like the reference, which never differentiates its features."""

    @staticmethod
    def forward(ctx, X, W):
        ctx.save_for_backward(X)
        return gemm(X, W)

    @staticmethod
    def backward(ctx, dS):
        (X,) = ctx.saved_tensors
        return None, gemm(X, dS.contiguous(), trans_a=True)


def _dropout_scale(p: float, training: bool) -> float:
    if not (training and p > 0.0):
        return 1.0
    return 1.0 / (1.0 - p) if p < 1.0 else 0.0  # p = 1 drops everything (torch.dropout accepts it): outputs and gradients are 0


class GCNCoreFunction(torch.autograd.Function):
    """logits = A @ (dropout(relu(A @ S1 + b1)) @ W2) + b2 as ONE autograd node (reference layer.py:181-188).

    forward : tg_gc1_fwd_f32 -> tg_dense_nn_f32 -> tg_spmm_f32(+b2)
    backward: tg_colsum_f32 (db2) -> tg_spmm_f32 on A^T (dS2) -> tg_hidden_bwd_f32 (dZ1, dW2, db1)
              -> tg_spmm_f32 on A^T (dS1)
    """

    @staticmethod
    def forward(ctx, S1, b1, W2, b2, csr: DeviceCSR, p: float, training: bool, keep_mask, seed: int, offset: int,
                offset_dev=None):
        H1 = gc1_forward(csr, S1, b1, p, training, keep_mask, seed, offset, offset_dev=offset_dev)
        S2 = dense_nn(H1, W2)
        logits = spmm(csr, S2, b2)
        ctx.save_for_backward(H1, W2)
        ctx.csr, ctx.scale = csr, _dropout_scale(p, training)
        ctx.has_b1, ctx.has_b2 = b1 is not None, b2 is not None
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        H1, W2 = ctx.saved_tensors
        dlogits = dlogits.contiguous()
        csr_t = ctx.csr.transpose()
        db2 = colsum(dlogits) if ctx.has_b2 else None
        dS2 = spmm(csr_t, dlogits)
        dZ1, dW2, db1 = hidden_backward(H1, dS2, W2, ctx.scale)
        dS1 = spmm(csr_t, dZ1) if ctx.needs_input_grad[0] else None
        return dS1, (db1 if ctx.has_b1 else None), dW2, db2, None, None, None, None, None, None, None


def identity_csr(n: int, device) -> DeviceCSR:
    """CSR of the n x n identity (cached): lets the fused layer-1 epilogue run on an already aggregated operand."""
    key = (n, str(device))
    hit = _identity_cache.get(key)
    if hit is None:
        rowptr = torch.arange(n + 1, dtype=torch.int32, device=device)
        colidx = torch.arange(n, dtype=torch.int32, device=device)
        vals = torch.ones(n, dtype=torch.float32, device=device)
        hit = DeviceCSR(rowptr, colidx, vals, n, n, symmetric=True)
        _identity_cache.clear()
        _identity_cache[key] = hit
    return hit


_identity_cache: dict = {}


class GCNLossFunction(torch.autograd.Function):
    """loss = mean_{train rows} CE(GCN(x, adj), y): the whole train-mode forward of trainer.py:357-359 with the
    log-softmax / NLL / index backward folded into the layer-2 epilogue (tg_gc2_loss_fwd_f32)."""

    @staticmethod
    def forward(ctx, S1, b1, W2, b2, csr: DeviceCSR, p: float, training: bool, keep_mask, seed: int, offset: int,
                row_label, inv_count: float, want_logits: bool, offset_dev=None):
        H1 = gc1_forward(csr, S1, b1, p, training, keep_mask, seed, offset, offset_dev=offset_dev)
        S2 = dense_nn(H1, W2)
        if callable(row_label):  # host labels: start their copy only now, with layer 1 already queued on the device
            row_label = row_label()
        loss, logits, dZ2 = gc2_loss_forward(csr, S2, b2, row_label, inv_count, want_logits=want_logits, want_grad=True)
        ctx.save_for_backward(H1, W2, dZ2)
        ctx.csr, ctx.scale = csr, _dropout_scale(p, training)
        ctx.has_b1, ctx.has_b2 = b1 is not None, b2 is not None
        if want_logits:
            ctx.mark_non_differentiable(logits)
            return loss, logits
        return loss, torch.empty(0, device=loss.device)

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        H1, W2, dZ2 = ctx.saved_tensors
        # dloss is a device scalar (1.0 for loss.backward()): it is folded into dS2 = (A^T dZ2) * dloss by the SpMM
        # epilogue and into db2 on 20 elements — no pass over the [N x C] gradient, no host sync
        g = dloss.reshape(()).to(torch.float32).contiguous()
        csr_t = ctx.csr.transpose()
        db2 = colsum(dZ2) * g if ctx.has_b2 else None
        dS2 = spmm(csr_t, dZ2, out_scale=g)
        dZ1, dW2, db1 = hidden_backward(H1, dS2, W2, ctx.scale)
        dS1 = spmm(csr_t, dZ1) if ctx.needs_input_grad[0] else None
        return dS1, (db1 if ctx.has_b1 else None), dW2, db2, None, None, None, None, None, None, None, None, None, None


class MaskedCrossEntropy(torch.autograd.Function):
    """mean CE over labelled rows of existing logits (reference trainer.py:358-359 semantics)."""

    @staticmethod
    def forward(ctx, logits, row_label, inv_count: float):
        loss, dZ = masked_ce(logits, row_label, inv_count, want_grad=True)
        ctx.save_for_backward(dZ)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dZ,) = ctx.saved_tensors
        return dZ * dloss, None, None


def masked_cross_entropy(logits: torch.Tensor, target: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """Drop-in for `CrossEntropyLoss()(logits[index], target[index])` (reference trainer.py:358-359).
    `target` holds one label per document (the first len(target) rows), `index` the rows that count."""
    row_label = make_row_label(int(logits.shape[0]), target, index)
    return MaskedCrossEntropy.apply(logits, row_label, 1.0 / max(int(index.numel()), 1))
