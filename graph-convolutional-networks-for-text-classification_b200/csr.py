"""Device CSR + skew plan for the normalised adjacency (and for sparse feature matrices).

The reference hands `torch.spmm` a COO tensor built by utils.sparse_mx_to_torch_sparse_tensor
(reference utils.py:196-203: int64 indices, fp32 values, flagged uncoalesced) and lets ATen re-sort and
re-convert it on every call (layer.py:106).  Here the tensor is converted ONCE to int32 CSR on the device
(tg_csr_from_coo), classified into short rows / hub rows (tg_plan_create), and cached per tensor.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import torch

from . import _native as N


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise N.TopicGCNError(
            f"{what} must live on a CUDA device (got {t.device}); topicgcn_b200 has no CPU fallback")


class DeviceCSR:
    """int32 CSR on the GPU + its SpMM plan.  `transpose()` returns the CSR used by the backward pass."""

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, vals: torch.Tensor, n_rows: int, n_cols: int,
                 *, hub_threshold: int = 0, segment_nnz: int = 0, symmetric: Optional[bool] = None,
                 streaming: bool = True):
        for t, name, dt in ((rowptr, "rowptr", torch.int32), (colidx, "colidx", torch.int32), (vals, "vals", torch.float32)):
            _require_cuda(t, name)
            if t.dtype != dt or not t.is_contiguous():
                raise N.TopicGCNError(f"{name} must be a contiguous {dt} tensor")
        if rowptr.numel() != n_rows + 1:
            raise N.TopicGCNError("rowptr must have n_rows + 1 entries")
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.nnz = int(colidx.numel())
        self.device = rowptr.device
        self._hub_threshold, self._segment_nnz = int(hub_threshold), int(segment_nnz)
        self._symmetric = symmetric
        self._transposed: Optional["DeviceCSR"] = None
        self._workspace: Optional[torch.Tensor] = None
        self._plan = C.c_void_p()
        self._streaming_requested = bool(streaming)
        with torch.cuda.device(self.device):
            # vals = NULL -> no column-chunk streaming layout (gather kernel only; colidx still orders its segments)
            N.check(N.lib().tg_plan_create(N.ptr(rowptr), N.ptr(colidx),
                                           N.ptr(vals) if streaming else 0, self.n_rows, self.n_cols, self.nnz,
                                           self._hub_threshold, self._segment_nnz, C.byref(self._plan),
                                           N.current_stream_ptr()), "tg_plan_create")
        info = (C.c_int64 * 8)()
        N.check(N.lib().tg_plan_info(self._plan, info), "tg_plan_info")
        self.n_hub_rows, self.n_segments, self.hub_nnz, self.max_row_nnz = (int(info[i]) for i in range(4))
        self.hub_threshold, self.segment_nnz = int(info[4]), int(info[5])
        self.streaming = bool(info[6] & 1)  # role-specialised streaming kernels (csrc/tg_roles2.cu) on the square graph
        self.roles2 = bool(info[6] & 2)
        self.roles2_rect = int(info[6] >> 2) & 3  # 1: resident-table product (X @ W), 2: all-hub product (X^T @ dS)
        self.chunk_rows = int(info[7]) & 0xFFFF
        self.hub_groups = (int(info[7]) >> 16) & 0xFF   # hub slot groups of 256 (K = 1024 topics: 4-5)
        self.doc_nq = (int(info[7]) >> 24) & 0xFF       # float4 chunks per lane of the document role (slice = 32 * nq columns)
        self.hub_gs = (int(info[7]) >> 32) & 0xFF       # lanes per hub slot sub-group (32 / 16 / 8: 256 / 512 / 1024 slots per group)

    def spmm_launches(self, B: torch.Tensor, n_feat: int, philox: bool = False, out_vec4_ok: bool = True,
                      loss: bool = False) -> int:
        """Kernels one tg_spmm* / tg_gc* call on this matrix launches (tg_plan_spmm_launches: the library's own kernel
        selection, no Python mirror of it).  philox: the call draws a Philox dropout mask; loss: it is the fused loss forward."""
        ld = int(B.stride(0)) if B.shape[0] > 1 else int(B.shape[1])
        mode = 2 if loss else int(bool(philox))
        return int(N.lib().tg_plan_spmm_launches(self._plan, N.ptr(B), ld, int(n_feat), mode, int(bool(out_vec4_ok))))

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None and self._plan.value:
                N.lib().tg_plan_destroy(self._plan)
                self._plan = C.c_void_p()
        except Exception:  # interpreter shutdown
            pass

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def from_coo(cls, rows: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, n_rows: int, n_cols: int,
                 **plan_kw) -> "DeviceCSR":
        """int64 COO triplets (any order, duplicates allowed) -> CSR with torch `coalesce()` semantics."""
        for t, name in ((rows, "rows"), (cols, "cols"), (vals, "vals")):
            _require_cuda(t, name)
        rows = rows.to(torch.int64).contiguous()
        cols = cols.to(torch.int64).contiguous()
        vals = vals.to(torch.float32).contiguous()
        nnz = int(rows.numel())
        dev = rows.device
        rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        colidx = torch.empty(nnz, dtype=torch.int32, device=dev)
        vout = torch.empty(nnz, dtype=torch.float32, device=dev)
        nnz_out, flags = C.c_int64(0), C.c_uint32(0)
        with torch.cuda.device(dev):
            N.check(N.lib().tg_csr_from_coo(N.ptr(rows), N.ptr(cols), N.ptr(vals), nnz, n_rows, n_cols, N.ptr(rowptr),
                                            N.ptr(colidx), N.ptr(vout), C.byref(nnz_out), C.byref(flags),
                                            N.current_stream_ptr()), "tg_csr_from_coo")
        m = int(nnz_out.value)
        if m != nnz:
            colidx, vout = colidx[:m].clone(), vout[:m].clone()
        out = cls(rowptr, colidx, vout, n_rows, n_cols, **plan_kw)
        out.coo_flags = int(flags.value)
        return out

    @classmethod
    def from_torch_sparse(cls, t: torch.Tensor, **plan_kw) -> "DeviceCSR":
        """torch.sparse COO (reference utils.py:203 layout) or torch sparse CSR tensor -> DeviceCSR."""
        if t.layout == torch.sparse_coo:
            _require_cuda(t, "sparse tensor")
            idx = t._indices()
            return cls.from_coo(idx[0], idx[1], t._values(), t.shape[0], t.shape[1], **plan_kw)
        if t.layout == torch.sparse_csr:
            _require_cuda(t, "sparse tensor")
            return cls(t.crow_indices().to(torch.int32).contiguous(), t.col_indices().to(torch.int32).contiguous(),
                       t.values().to(torch.float32).contiguous(), t.shape[0], t.shape[1], **plan_kw)
        raise N.TopicGCNError(f"unsupported sparse layout {t.layout}")

    # ------------------------------------------------------------------------------------------------
    def transpose(self) -> "DeviceCSR":
        """CSR of A^T.  The normalised adjacency is exactly symmetric, in which case the
        same arrays are reused and no second copy is kept."""
        if self._symmetric:
            return self
        if self._transposed is None:
            dev = self.device
            t_rowptr = torch.empty(self.n_cols + 1, dtype=torch.int32, device=dev)
            t_colidx = torch.empty(self.nnz, dtype=torch.int32, device=dev)
            t_vals = torch.empty(self.nnz, dtype=torch.float32, device=dev)
            sym = C.c_int32(0)
            with torch.cuda.device(dev):
                N.check(N.lib().tg_csr_transpose(N.ptr(self.rowptr), N.ptr(self.colidx), N.ptr(self.vals), self.n_rows,
                                                 self.n_cols, self.nnz, N.ptr(t_rowptr), N.ptr(t_colidx),
                                                 N.ptr(t_vals), C.byref(sym), N.current_stream_ptr()),
                        "tg_csr_transpose")
            if sym.value == 1:
                self._symmetric = True
                return self
            self._symmetric = False
            self._transposed = DeviceCSR(t_rowptr, t_colidx, t_vals, self.n_cols, self.n_rows,
                                         hub_threshold=self._hub_threshold, segment_nnz=self._segment_nnz,
                                         symmetric=False, streaming=self._streaming_requested)
            self._transposed._transposed = self
        return self._transposed

    @property
    def is_symmetric(self) -> bool:
        self.transpose()
        return bool(self._symmetric)

    def workspace(self, n_feat: int):
        """(ptr, bytes) of the scratch the split hub rows need for `n_feat` columns (grown on demand)."""
        need = int(N.lib().tg_plan_workspace_bytes(self._plan, int(n_feat)))
        if need == 0:
            return 0, 0
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace.data_ptr(), need

    @property
    def plan(self):
        return self._plan

    def to_torch_coo(self) -> torch.Tensor:
        """Back to a coalesced torch COO tensor (tests)."""
        counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), counts)
        idx = torch.stack([rows, self.colidx.to(torch.int64)])
        return torch.sparse_coo_tensor(idx, self.vals, (self.n_rows, self.n_cols)).coalesce()


# ---- per-tensor cache: the trainer passes the SAME adj / feature tensors every epoch (trainer.py:357,382) ----
_cache: dict = {}


def _cache_key(t: torch.Tensor):
    # identity + storage + version counters: an in-place edit of the values or indices (same storage) bumps `_version`, so the
    # next call converts again instead of serving the stale CSR (torch.spmm would see the new values too)
    if t.layout == torch.sparse_coo:
        v, i = t._values(), t._indices()
        return (id(t), v.data_ptr(), i.data_ptr(), v._version, i._version, t._nnz(), tuple(t.shape))
    v, c = t.values(), t.col_indices()
    return (id(t), v.data_ptr(), c.data_ptr(), v._version, c._version, tuple(t.shape))


def cached_csr(t: torch.Tensor, **plan_kw) -> DeviceCSR:
    """DeviceCSR of a torch sparse tensor, converted once and remembered for as long as the tensor lives."""
    key = _cache_key(t)
    hit = _cache.get(key)
    if hit is not None and hit[0]() is t:
        return hit[1]
    csr = DeviceCSR.from_torch_sparse(t, **plan_kw)
    _cache[key] = (weakref.ref(t, lambda _r, k=key: _cache.pop(k, None)), csr)
    return csr


def clear_cache() -> None:
    _cache.clear()
