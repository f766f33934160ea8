"""CPU oracle of the TopicGCN hot path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in numpy + the C loop of oracle/spmm_oracle.c, what the reference computes on the path:

  normalize_adj / preprocess_adj      reference utils.py:185-213   -> normalize_adj_coo
  sparse_mx_to_torch_sparse_tensor    reference utils.py:196-203   -> (rows int64, cols int64, vals fp32) triplets
  GraphConvolution.forward            reference layer.py:84-112    -> graph_convolution
  GCN.forward                         reference layer.py:164-190   -> gcn_forward
  CrossEntropyLoss on train rows      reference trainer.py:358-359 -> masked_cross_entropy
  loss.backward()                     autograd graph, SURVEY §3.3  -> gcn_loss_and_grads
  Adam step                           reference trainer.py:307,362 -> adam_step (torch.optim.Adam defaults)

The sparse x dense product itself is third-party code (PyTorch ATen `s_addmm_out_sparse_dense_worker`, torch pinned
1.6.0 in the reference's requirements.txt:2): an fp32 axpy per stored entry in storage order — restated in C.

PARITY PINNING: the reference has no golden vectors for this path.  tests/golden/*.npz were produced by importing
the REAL reference modules (`layer.GCN`, `utils.preprocess_adj`) from /root/reference in the build container with
tests/golden/make_golden.py; tests/test_oracle.py checks every function here against them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/spmm_oracle.c with gcc (Makefile in this directory)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "spmm_oracle.c")):
        subprocess.run(["make", "-B", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def _c():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_csr_from_coo.restype = C.c_int64
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(_c().oracle_num_threads())


# ---------------------------------------------------------------------------------------------------------------
# sparse containers
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Coo:
    """The torch.sparse COO tensor of reference utils.py:203 as plain arrays."""
    rows: np.ndarray  # int64 [nnz]
    cols: np.ndarray  # int64 [nnz]
    vals: np.ndarray  # float32 [nnz]
    shape: tuple

    def __post_init__(self):
        self.rows = np.ascontiguousarray(self.rows, dtype=np.int64)
        self.cols = np.ascontiguousarray(self.cols, dtype=np.int64)
        self.vals = np.ascontiguousarray(self.vals, dtype=np.float32)

    @property
    def nnz(self) -> int:
        return int(self.rows.size)

    def transpose(self) -> "Coo":
        """`sparse.t()`: swaps the index rows, storage order unchanged; memoised."""
        t = getattr(self, "_t", None)
        if t is None:
            t = Coo(self.cols, self.rows, self.vals, (self.shape[1], self.shape[0]))
            self._t = t
        return t


def csr_from_coo(coo: Coo):
    """torch `coalesce()` + row pointer: (rowptr int32, colidx int32, vals fp32)."""
    n_rows, n_cols = coo.shape
    rowptr = np.zeros(n_rows + 1, dtype=np.int32)
    colidx = np.zeros(max(coo.nnz, 1), dtype=np.int32)
    vals = np.zeros(max(coo.nnz, 1), dtype=np.float32)
    m = _c().oracle_csr_from_coo(_ptr(coo.rows), _ptr(coo.cols), _ptr(coo.vals), C.c_int64(coo.nnz),
                                 C.c_int64(n_rows), C.c_int64(n_cols), _ptr(rowptr), _ptr(colidx), _ptr(vals))
    return rowptr, colidx[:m].copy(), vals[:m].copy()


# ---------------------------------------------------------------------------------------------------------------
# the sparse x dense product (layer.py:102/106 -> ATen)
# ---------------------------------------------------------------------------------------------------------------
_THREADS = 1


def set_threads(n: int) -> None:
    """n > 1: `spmm` runs the row-parallel CSR form of the same loop on n OpenMP threads (bit-identical results for
    any input: rows are gathered with a stable sort).  Used by the CPU-baseline legs of bench.py; the default (1) is the reference's serial
    storage-order loop."""
    global _THREADS
    _THREADS = max(1, int(n))


def _row_sorted_csr(coo: "Coo"):
    hit = getattr(coo, "_csr_cache", None)
    if hit is None:
        # a STABLE sort by row keeps, inside every row, the storage order of its entries: the row-parallel loop
        # then performs exactly the additions of the serial storage-order loop, row by row
        cols, vals = coo.cols, coo.vals
        if coo.nnz > 1 and not bool(np.all(coo.rows[1:] >= coo.rows[:-1])):
            order = np.argsort(coo.rows, kind="stable")
            cols, vals = cols[order], vals[order]
        rowptr = np.zeros(coo.shape[0] + 1, dtype=np.int32)
        np.cumsum(np.bincount(coo.rows, minlength=coo.shape[0]), out=rowptr[1:])
        hit = (rowptr, cols.astype(np.int32), np.ascontiguousarray(vals))
        coo._csr_cache = hit
    return hit


def spmm(coo: Coo, B: np.ndarray) -> np.ndarray:
    """th.spmm(sparse_coo, dense) on CPU: serial fp32 axpy per stored entry, storage order."""
    B = np.ascontiguousarray(B, dtype=np.float32)
    F = B.shape[1]
    if _THREADS > 1:
        csr = _row_sorted_csr(coo)
        return spmm_csr(csr[0], csr[1], csr[2], B, _THREADS)
    Y = np.empty((coo.shape[0], F), dtype=np.float32)
    _c().oracle_spmm_coo_f32(_ptr(coo.rows), _ptr(coo.cols), _ptr(coo.vals), C.c_int64(coo.nnz), _ptr(B),
                             C.c_int64(F), C.c_int64(F), _ptr(Y), C.c_int64(F), C.c_int64(coo.shape[0]))
    return Y


def spmm_f64(coo: Coo, B: np.ndarray) -> np.ndarray:
    """Same sums with fp64 accumulation (ground truth for error comparisons)."""
    B = np.ascontiguousarray(B, dtype=np.float32)
    F = B.shape[1]
    Y = np.empty((coo.shape[0], F), dtype=np.float64)
    _c().oracle_spmm_coo_f64(_ptr(coo.rows), _ptr(coo.cols), _ptr(coo.vals), C.c_int64(coo.nnz), _ptr(B),
                             C.c_int64(F), C.c_int64(F), _ptr(Y), C.c_int64(F), C.c_int64(coo.shape[0]))
    return Y


def row_dot_f32(vals: np.ndarray, rows_of_B: np.ndarray) -> np.ndarray:
    """One output row of `spmm`: sum_p vals[p] * rows_of_B[p] accumulated serially in fp32, storage order (the same loop,
    run on a one-row matrix)."""
    m = int(vals.size)
    one = Coo(np.zeros(m, dtype=np.int64), np.arange(m, dtype=np.int64), np.asarray(vals, dtype=np.float32), (1, max(m, 1)))
    if m == 0:
        return np.zeros(rows_of_B.shape[1], dtype=np.float32)
    return spmm(one, rows_of_B)[0]


def spmm_csr(rowptr, colidx, vals, B: np.ndarray, n_threads: int = 0) -> np.ndarray:
    """Row-parallel form of the same loop (bit-identical for row-major sorted input); the CPU baseline."""
    B = np.ascontiguousarray(B, dtype=np.float32)
    F = B.shape[1]
    n_rows = rowptr.size - 1
    Y = np.empty((n_rows, F), dtype=np.float32)
    _c().oracle_spmm_csr_f32(_ptr(rowptr), _ptr(colidx), _ptr(vals), C.c_int64(n_rows), _ptr(B), C.c_int64(F),
                             C.c_int64(F), _ptr(Y), C.c_int64(F), C.c_int(n_threads))
    return Y


# ---------------------------------------------------------------------------------------------------------------
# adjacency normalisation (utils.py:185-213)
# ---------------------------------------------------------------------------------------------------------------
def normalize_adj_coo(rows, cols, vals, n: int) -> Coo:
    """Â = ((A+I) D^-1/2)^T D^-1/2 computed in float64 and cast to fp32, entries in row-major order — the exact
    output of utils.preprocess_adj(adj, is_sparse=True) for a duplicate-free input (rows, cols, vals).

      (A+I)            utils.py:188   (sp.eye is float64, so everything below is float64)
      rowsum           utils.py:209   sum over each row in column order
      d = rowsum^-1/2  utils.py:210-211, inf -> 0
      Â[i,j] = (Ã[j,i] * d[i]) * d[j]   utils.py:213 (adj.dot(D).transpose().dot(D)), then .astype(float32)
    """
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    v = np.asarray(vals).astype(np.float64)
    # A + I, merged and sorted row-major
    r = np.concatenate([rows, np.arange(n, dtype=np.int64)])
    c = np.concatenate([cols, np.arange(n, dtype=np.int64)])
    v = np.concatenate([v, np.ones(n, dtype=np.float64)])
    key = r * n + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    head = np.ones(key.size, dtype=bool)
    head[1:] = key[1:] != key[:-1]
    seg = np.cumsum(head) - 1
    vv = np.zeros(int(seg[-1]) + 1, dtype=np.float64)
    np.add.at(vv, seg, v)  # explicit diagonal entries of A are added to the 1 of I
    key = key[head]
    r, c = key // n, key % n
    rowsum = np.bincount(r, weights=vv, minlength=n)  # sequential, column order within a row
    with np.errstate(divide="ignore"):
        d = np.power(rowsum, -0.5)
    d[np.isinf(d)] = 0.0
    # entry (i, j) of the result takes Ã[j, i]: walk the transposed matrix in row-major order
    tkey = c * n + r
    torder = np.argsort(tkey, kind="stable")
    ti, tj, tv = c[torder], r[torder], vv[torder]  # result row i = old col, result col j = old row
    out = (tv * d[ti]) * d[tj]
    return Coo(ti, tj, out.astype(np.float32), (n, n))


# ---------------------------------------------------------------------------------------------------------------
# dropout masks
# ---------------------------------------------------------------------------------------------------------------
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(_W0)) & mask
        k1 = (k1 + np.uint64(_W1)) & mask
    return c0, c1, c2, c3


def philox_keep_mask(n_rows: int, n_feat: int, p: float, seed: int, offset: int) -> np.ndarray:
    """The counter-based keep mask of the CUDA library (definition in csrc/tg_common.cuh), restated in numpy."""
    return philox_keep_mask_rows(np.arange(n_rows, dtype=np.uint64), n_feat, p, seed, offset)


def philox_keep_mask_rows(row_ids, n_feat: int, p: float, seed: int, offset: int) -> np.ndarray:
    """The keep mask of the given rows only (the mask is a pure function of (row, column, seed, offset))."""
    row_ids = np.asarray(row_ids, dtype=np.uint64)
    n_rows = int(row_ids.size)
    thr = int(min(max((1.0 - np.float32(p)) * np.float32(65536.0) + np.float32(0.5), 0.0), 65536.0))
    row = np.repeat(row_ids, n_feat)
    col = np.tile(np.arange(n_feat, dtype=np.uint64), n_rows)
    q = col >> np.uint64(2)
    slot, j, half = q & np.uint64(7), q >> np.uint64(4), (q >> np.uint64(3)) & np.uint64(1)
    seed, offset = int(seed) & (2**64 - 1), int(offset) & (2**64 - 1)
    k0 = (seed & 0xFFFFFFFF) ^ (offset >> 32)
    k1 = seed >> 32
    if thr == 32768:
        # exact-half mode (p = 0.5): one random bit per element, one call per 128 columns (csrc/tg_common.cuh)
        cidx = (col >> np.uint64(7)) | np.uint64(0x80000000)
        x, y, z, w = _philox4x32_10(row & np.uint64(0xFFFFFFFF), row >> np.uint64(32), cidx,
                                    np.full(row.shape, offset & 0xFFFFFFFF, dtype=np.uint64), k0, k1)
        wsel = (col >> np.uint64(5)) & np.uint64(3)
        word = np.where(wsel == 0, x, np.where(wsel == 1, y, np.where(wsel == 2, z, w)))
        bit = (word >> (col & np.uint64(31))) & np.uint64(1)
        return bit.astype(np.uint8).reshape(n_rows, n_feat)
    x, y, z, w = _philox4x32_10(row & np.uint64(0xFFFFFFFF), row >> np.uint64(32), slot | (j << np.uint64(8)),
                                np.full(row.shape, offset & 0xFFFFFFFF, dtype=np.uint64), k0, k1)
    a = np.where(half == 1, z, x)
    b = np.where(half == 1, w, y)
    e = col & np.uint64(3)
    word = np.where(e < 2, a, b)
    u16 = np.where((e & np.uint64(1)) == 1, word >> np.uint64(16), word & np.uint64(0xFFFF))
    return (u16 < np.uint64(thr)).astype(np.uint8).reshape(n_rows, n_feat)


# ---------------------------------------------------------------------------------------------------------------
# the model (layer.py) and the loss (trainer.py)
# ---------------------------------------------------------------------------------------------------------------
def _support(x, W: np.ndarray) -> np.ndarray:
    """support = th.spmm(infeatn, self.weight)  (layer.py:102); x is None for the featureless identity."""
    if x is None:
        return W
    if isinstance(x, Coo):
        return spmm(x, W)
    return (np.asarray(x, dtype=np.float32) @ W).astype(np.float32)


def graph_convolution(x, adj: Coo, W: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    """GraphConvolution.forward (layer.py:84-112)."""
    out = spmm(adj, _support(x, W))
    return out + b if b is not None else out


def gcn_forward(x, adj: Coo, params: dict, p: float = 0.5, training: bool = False,
                keep_mask: Optional[np.ndarray] = None):
    """GCN.forward (layer.py:164-190).  Returns (logits, cache).  In training mode `keep_mask` is the [N x nhid]
    Bernoulli(1-p) sample the reference draws inside th.dropout (layer.py:185)."""
    W1, b1, W2, b2 = (np.asarray(params[k], dtype=np.float32) for k in ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"))
    S1 = _support(x, W1)
    Z1 = spmm(adj, S1) + b1                                  # layer.py:181 (gc1)
    A1 = np.maximum(Z1, np.float32(0))                       # layer.py:182
    if training and p > 0:
        scale = np.float32(1.0 / (1.0 - p))
        H1 = (A1 * (keep_mask.astype(np.float32) * scale)).astype(np.float32)   # layer.py:185: noise.div_(1-p); x*noise
    else:
        scale = np.float32(1.0)
        H1 = A1
    S2 = (H1 @ W2).astype(np.float32)                        # layer.py:188 -> :102 (dense)
    logits = spmm(adj, S2) + b2                              # layer.py:188 -> :106,110
    return logits, dict(S1=S1, Z1=Z1, H1=H1, S2=S2, scale=scale, keep_mask=keep_mask if training and p > 0 else None)


def masked_cross_entropy(logits: np.ndarray, target: np.ndarray, index: np.ndarray):
    """CrossEntropyLoss()(logits[index], target[index]) (trainer.py:358-359): (loss, dlogits [N x C])."""
    z = logits[index].astype(np.float32)
    y = np.asarray(target)[index].astype(np.int64)
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m, dtype=np.float32)
    lse = m[:, 0] + np.log(e.sum(axis=1, dtype=np.float32), dtype=np.float32)
    T = index.size
    loss = np.float32((lse - z[np.arange(T), y]).sum(dtype=np.float64) / T)
    sm = np.exp(z - lse[:, None], dtype=np.float32)
    sm[np.arange(T), y] -= np.float32(1)
    dlogits = np.zeros_like(logits, dtype=np.float32)
    dlogits[index] = sm / np.float32(T)
    return loss, dlogits


def gcn_loss_and_grads(x, adj: Coo, params: dict, target, index, p: float = 0.5, training: bool = True,
                       keep_mask: Optional[np.ndarray] = None):
    """One train-mode forward + loss + backward (trainer.py:357-361).  Gradient flow restates the autograd graph of
    SURVEY §3.3: each sparse MmBackward is `sparse.t().mm(grad)`."""
    logits, cache = gcn_forward(x, adj, params, p, training, keep_mask)
    loss, dZ2 = masked_cross_entropy(logits, np.asarray(target), np.asarray(index))
    W2 = np.asarray(params["gc2.weight"], dtype=np.float32)
    adj_t = adj.transpose()
    grads = {}
    grads["gc2.bias"] = dZ2.sum(axis=0, dtype=np.float32)
    dS2 = spmm(adj_t, dZ2)
    grads["gc2.weight"] = (cache["H1"].T @ dS2).astype(np.float32)
    dH1 = (dS2 @ W2.T).astype(np.float32)
    if cache["keep_mask"] is not None:
        dH1 = dH1 * (cache["keep_mask"].astype(np.float32) * cache["scale"])
    dZ1 = np.where(cache["Z1"] > 0, dH1, np.float32(0)).astype(np.float32)
    grads["gc1.bias"] = dZ1.sum(axis=0, dtype=np.float32)
    dS1 = spmm(adj_t, dZ1)
    if x is None:
        grads["gc1.weight"] = dS1
    elif isinstance(x, Coo):
        grads["gc1.weight"] = spmm(x.transpose(), dS1)
    else:
        grads["gc1.weight"] = (np.asarray(x, dtype=np.float32).T @ dS1).astype(np.float32)
    return loss, logits, grads


def adam_step(params: dict, grads: dict, state: dict, lr: float = 0.02, b1: float = 0.9, b2: float = 0.999,
              eps: float = 1e-8) -> None:
    """torch.optim.Adam with default betas/eps (trainer.py:307), in place."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for k in params:
        g = grads[k].astype(np.float32)
        m = state.setdefault("m_" + k, np.zeros_like(g))
        v = state.setdefault("v_" + k, np.zeros_like(g))
        m *= np.float32(b1); m += np.float32(1 - b1) * g
        v *= np.float32(b2); v += np.float32(1 - b2) * g * g
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        step = lr / bc1
        denom = np.sqrt(v) / np.float32(np.sqrt(bc2)) + np.float32(eps)
        params[k] -= (np.float32(step) * m / denom).astype(np.float32)


def class_counts(logits: np.ndarray, target: np.ndarray, index: np.ndarray, n_class: int) -> np.ndarray:
    """[3 x n_class] true positives / false positives / false negatives of argmax(logits[index]) against target[index]
    (the per-class sums of utils.macro_f1, reference utils.py:57-69; ties resolve to the lowest class like th.max)."""
    pred = logits[index].argmax(axis=1)
    targ = np.asarray(target)[index]
    out = np.zeros((3, n_class), dtype=np.int64)
    for c in range(n_class):
        out[0, c] = np.sum((pred == c) & (targ == c))
        out[1, c] = np.sum((pred == c) & (targ != c))
        out[2, c] = np.sum((pred != c) & (targ == c))
    return out


def metrics_from_counts(counts: np.ndarray, n_rows: int) -> dict:
    """accuracy (utils.py:89-109) and macro F1 / precision / recall (utils.py:71-85: per-class ratios with 0/0 -> 0,
    macro-averaged, F1 of the two averages) from the per-class counts."""
    tp, fp, fn = (np.asarray(counts[i], dtype=np.float64) for i in range(3))
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = tp / (tp + fp)
        rec = tp / (tp + fn)
    prec[np.isnan(prec)] = 0
    rec[np.isnan(rec)] = 0
    P, R = float(prec.mean()), float(rec.mean())
    with np.errstate(divide="ignore", invalid="ignore"):
        f1 = float(np.float64(2 * P * R) / np.float64(P + R))
    return {"acc": float(tp.sum()) / max(int(n_rows), 1), "macro_f1": f1, "precision": P, "recall": R}


def accuracy(logits: np.ndarray, target: np.ndarray, index: np.ndarray) -> float:
    """utils.accuracy (reference utils.py:89-109) on logits[index]."""
    return float((logits[index].argmax(axis=1) == np.asarray(target)[index]).mean())
