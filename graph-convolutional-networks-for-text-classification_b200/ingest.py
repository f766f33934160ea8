"""Edge-list ingest without networkx / scipy: the graph file the reference's builder writes
(`nx.write_weighted_edgelist`, build_graph.py:199: one `u v w` line per undirected edge, integer node ids) straight to the
normalised adjacency `Â = D^-1/2 (A + I) D^-1/2` the model consumes — the same tensor, bit for bit, that the reference
makes with `nx.read_weighted_edgelist(nodetype=int)` -> `nx.adjacency_matrix(nodelist=range(n), dtype=float32)` ->
symmetrise -> `utils.preprocess_adj` (trainer.py:98-151, utils.py:185-213), but with the sort / degree / scaling work done
by tensor ops on the target device (`graphgen.normalize_undirected`; 7.9 s through networkx + scipy at 1 M documents).

Semantics kept from the reference:
  * node ids are matrix indices; n = number of distinct ids, and every id in 0..n-1 must occur (the reference's
    `nodelist=range(n)` raises otherwise — so does this);
  * an edge that occurs more than once (in either orientation) keeps its LAST weight (`nx.Graph.add_edge` overwrites);
  * weights are parsed as float64 and rounded to fp32 (`dtype=np.float32`); a self loop `u u w` adds w to the diagonal
    next to the identity of A + I.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N
from . import graphgen


def read_edge_list(path: str):
    """(u int64, v int64, w float32, n) of an `u v w` text file with the reference's duplicate / orientation rules."""
    import pandas as pd
    df = pd.read_csv(path, sep=r"\s+", header=None, names=["u", "v", "w"], comment="#",
                     dtype={"u": np.int64, "v": np.int64, "w": np.float64}, engine="c")
    u, v = df["u"].to_numpy(), df["v"].to_numpy()
    w = df["w"].to_numpy().astype(np.float32)  # nx.adjacency_matrix(dtype=np.float32)
    if u.size and (min(u.min(), v.min()) < 0):
        raise N.TopicGCNError("negative node id in the edge list")
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    ids = np.unique(np.concatenate([u, v]))
    n = int(ids.size)
    if n and int(ids[-1]) != n - 1:
        raise N.TopicGCNError(f"edge list names {n} distinct nodes but ids reach {int(ids[-1])}: every id in 0..n-1 must "
                              "occur (the reference's nodelist=range(n) fails the same way)")
    # the last occurrence of an undirected pair wins
    key = lo * max(n, 1) + hi
    _, last = np.unique(key[::-1], return_index=True)
    keep = np.sort(key.size - 1 - last)
    return lo[keep], hi[keep], w[keep], n


def normalized_adjacency(u, v, w, n: int, device="cuda"):
    """Row-major COO (rows int64, cols int64, vals fp32) of Â for unique undirected edges given as numpy arrays."""
    dev = torch.device(device)
    u = torch.as_tensor(u, dtype=torch.int64, device=dev)
    v = torch.as_tensor(v, dtype=torch.int64, device=dev)
    w = torch.as_tensor(w, dtype=torch.float32, device=dev)
    loop = u == v
    diag = None
    if bool(loop.any()):
        diag = torch.zeros(n, dtype=torch.float64, device=dev).index_add_(0, u[loop], w[loop].to(torch.float64))
        u, v, w = u[~loop], v[~loop], w[~loop]
    return graphgen.normalize_undirected(u, v, w, n, diag_extra=diag)


def load_adjacency(path: str, device="cuda") -> torch.Tensor:
    """The torch.sparse COO tensor the reference hands to the model (utils.py:203; not flagged coalesced)."""
    u, v, w, n = read_edge_list(path)
    rows, cols, vals = normalized_adjacency(u, v, w, n, device)
    return torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (n, n), check_invariants=False)
