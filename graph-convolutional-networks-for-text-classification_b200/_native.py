"""ctypes binding of libtopicgcn.so (include/topicgcn.h).

This is the ONLY compute backend of the package: there is no CPU or eager-PyTorch fallback.  If the shared
library is missing, or a call returns a non-zero status, a TopicGCNError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libtopicgcn.so")

# every symbol include/topicgcn.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "tg_version", "tg_last_error", "tg_status_string",
    "tg_csr_from_coo", "tg_csr_transpose",
    "tg_plan_create", "tg_plan_destroy", "tg_plan_info", "tg_plan_workspace_bytes", "tg_plan_spmm_launches",
    "tg_spmm_f32", "tg_gc1_fwd_f32", "tg_dropout_keep_mask", "tg_gc2_loss_fwd_f32", "tg_masked_ce_f32",
    "tg_reduce_scratch_floats", "tg_reduce_sum_f32",
    "tg_dense_nn_f32", "tg_hidden_bwd_scratch_floats", "tg_hidden_bwd_f32", "tg_hidden_bwd_rows_f32",
    "tg_colsum_scratch_floats", "tg_colsum_f32", "tg_relu_dropout_bwd_f32", "tg_adam_f32", "tg_class_counts_i32",
    "tg_gemm_scratch_floats", "tg_gemm_f32",
]


class TopicGCNError(RuntimeError):
    pass


_lib = None

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f32 = C.c_float
_u64 = C.c_uint64
_sz = C.c_size_t


def _declare(lib) -> None:
    def sig(name, restype, *argtypes):
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    sig("tg_version", C.c_int)
    sig("tg_last_error", C.c_char_p)
    sig("tg_status_string", C.c_char_p, C.c_int)
    sig("tg_csr_from_coo", C.c_int, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, C.POINTER(_i64),
        C.POINTER(C.c_uint32), _p)
    sig("tg_csr_transpose", C.c_int, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, C.POINTER(_i32), _p)
    sig("tg_plan_create", C.c_int, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, C.POINTER(_p), _p)
    sig("tg_plan_destroy", None, _p)
    sig("tg_plan_info", C.c_int, _p, C.POINTER(_i64))
    sig("tg_plan_workspace_bytes", _sz, _p, _i32)
    sig("tg_plan_spmm_launches", C.c_int, _p, _p, _i64, _i32, _i32, _i32)
    sig("tg_spmm_f32", C.c_int, _p, _p, _p, _p, _p, _i64, _p, _i64, _i32, _p, _p, _p, _sz, _p)
    sig("tg_gc1_fwd_f32", C.c_int, _p, _p, _p, _p, _p, _i64, _p, _p, _i64, _i32, _f32, _i32, _p, _u64, _u64,
        _p, _i64, _p, _sz, _p)
    sig("tg_dropout_keep_mask", C.c_int, _p, _i64, _i32, _f32, _u64, _u64, _p)
    sig("tg_gc2_loss_fwd_f32", C.c_int, _p, _p, _p, _p, _p, _i64, _p, _p, _f32, _p, _i64, _p, _i64, _p, _i32,
        _p, _sz, _p)
    sig("tg_masked_ce_f32", C.c_int, _p, _i64, _p, _f32, _p, _i64, _p, _i64, _i32, _p)
    sig("tg_reduce_scratch_floats", _i64, _i64)
    sig("tg_reduce_sum_f32", C.c_int, _p, _i64, _p, _p, _p)
    sig("tg_dense_nn_f32", C.c_int, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p)
    sig("tg_hidden_bwd_scratch_floats", _i64, _i64, _i32, _i32)
    sig("tg_hidden_bwd_f32", C.c_int, _p, _i64, _p, _i64, _p, _i64, _f32, _p, _i64, _p, _p, _p, _i64, _i32,
        _i32, _p)
    sig("tg_hidden_bwd_rows_f32", C.c_int, _p, _i64, _p, _i64, _p, _i64, _f32, _p, _i64, _p, _p, _p, _i64, _i32,
        _i32, _i64, _p)
    sig("tg_colsum_scratch_floats", _i64, _i64, _i32)
    sig("tg_colsum_f32", C.c_int, _p, _i64, _i64, _i32, _p, _p, _p)
    sig("tg_relu_dropout_bwd_f32", C.c_int, _p, _i64, _p, _i64, _f32, _p, _i64, _i64, _i32, _p)
    sig("tg_class_counts_i32", C.c_int, _p, _i64, _p, _i64, _i32, _p, _p)
    sig("tg_gemm_scratch_floats", _i64, _i32, _i64, _i64, _i64)
    sig("tg_gemm_f32", C.c_int, _i32, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _p, _p)
    sig("tg_adam_f32", C.c_int, _p, _p, _p, _p, _i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _i64, _p)


def lib():
    """Load libtopicgcn.so (once).  Fails loudly: the CUDA library is the product, not an accelerator."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TopicGCNError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  topicgcn_b200 has no CPU / PyTorch fallback.")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as exc:  # pragma: no cover - depends on the host
            raise TopicGCNError(f"cannot load {LIB_PATH}: {exc}") from exc
        _declare(handle)
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        l = lib()
        msg = l.tg_last_error().decode(errors="replace")
        kind = l.tg_status_string(status).decode()
        raise TopicGCNError(f"{what} failed: {kind} ({status}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
