"""Build libtopicgcn.so (the C-ABI of include/topicgcn.h) in-tree with nvcc for sm_100a.

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.  No JIT, no
torch.utils.cpp_extension: the library has no torch types in its interface (plain C-ABI, loaded by ctypes).
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libtopicgcn.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")
SOURCES = ["tg_api.cu", "tg_csr.cu", "tg_spmm.cu", "tg_dense.cu", "tg_roles2.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the topicgcn CUDA library cannot be built")


def _sources() -> list[str]:
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint(src: str | None = None) -> str:
    """Hash of the flags, the C header, every .cuh and either one .cu (per-object stamp) or all of them (library stamp)."""
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh") or (f.endswith(".cu") and (src is None or f == src)))
    for f in files + [os.path.join(INCLUDE, "topicgcn.h")]:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: str, log_dir: str, force: bool = False) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    stamp = obj + ".stamp"
    fp = _fingerprint(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == fp:
        return obj  # this object is current: only changed translation units are recompiled
    cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(log_dir, src + ".ptxas.log"), "w") as fh:
        fh.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fh:
        fh.write(fp)
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every CUDA source for sm_100a and link libtopicgcn.so; returns its path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "fingerprint.txt")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == fp:
        if verbose:
            print(f"[topicgcn build] up to date: {LIB_PATH}")
        return LIB_PATH
    nvcc = _nvcc()
    srcs = _sources()
    if verbose:
        print(f"[topicgcn build] nvcc sm_100a: {', '.join(srcs)}")
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, OBJ_DIR, force), srcs))
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fh:
        fh.write(fp)
    if verbose:
        print(f"[topicgcn build] wrote {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
