"""Synthetic document-topic-topic graphs of the BASELINE.json shapes, built and normalised with tensor ops on
whatever device is asked for (CUDA for the benchmark sizes, CPU in the host-side tests).

Structure follows the reference's graph builder:
  doc-topic edges   weight theta_dk, kept when >= 0.02                (reference build_graph.py:99-114)
  topic-topic edges cosine similarity of topic embeddings, kept > 0.3, i < j   (build_graph.py:116-133)
and the reference's ingest + normalisation:
  undirected graph -> symmetric A (trainer.py:98-148) -> Â = ((A+I) D^-1/2)^T D^-1/2 in float64, cast to fp32,
  stored row-major COO with int64 indices (utils.py:185-213).

Bit-exactness with utils.preprocess_adj (tested against the oracle restatement and, through it, the real reference):
row sums are taken in float64 over fp32 weights bounded below by 0.02, so every partial sum is exactly representable
and the summation order is immaterial; d = rowsum^-0.5 is evaluated by numpy on the host exactly like the reference
(np.power, utils.py:210); the two float64 multiplications and the final fp32 rounding are IEEE operations.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch


@dataclass
class SyntheticGraph:
    n_docs: int
    n_hubs: int                     # topic (or word) nodes, numbered after the documents
    rows: torch.Tensor              # int64 [nnz]  row-major sorted COO of Â
    cols: torch.Tensor              # int64 [nnz]
    vals: torch.Tensor              # fp32  [nnz]
    labels: torch.Tensor            # int64 [n_docs]
    train_idx: torch.Tensor         # int64 sorted
    val_idx: torch.Tensor
    test_idx: torch.Tensor
    n_class: int
    meta: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return self.n_docs + self.n_hubs

    @property
    def nnz(self) -> int:
        return int(self.rows.numel())

    def adj(self) -> torch.Tensor:
        """The torch.sparse COO tensor the reference hands to the model (utils.py:203): not flagged coalesced."""
        return torch.sparse_coo_tensor(torch.stack([self.rows, self.cols]), self.vals, (self.n, self.n),
                                       check_invariants=False)


def normalize_undirected(u: torch.Tensor, v: torch.Tensor, w: torch.Tensor, n: int, diag_extra=None):
    """Unique undirected edges (u != v) with fp32 weights -> row-major COO of Â = D^-1/2 (A+I) D^-1/2 with the
    reference's arithmetic (utils.py:206-213).  Returns (rows int64, cols int64, vals fp32).
    diag_extra: optional float64 [n] added to the diagonal of A + I (self loops of an ingested edge list)."""
    dev = u.device
    ar = torch.arange(n, dtype=torch.int64, device=dev)
    r = torch.cat([u, v, ar])
    c = torch.cat([v, u, ar])
    a = torch.cat([w, w, torch.ones(n, dtype=torch.float32, device=dev)]).to(torch.float64)
    if diag_extra is not None:
        a[2 * u.numel():] += diag_extra.to(torch.float64)
    del ar
    order = torch.argsort(r * n + c)  # keys are unique: any sort is a row-major ordering
    r, c, a = r[order], c[order], a[order]
    del order
    rowsum = torch.zeros(n, dtype=torch.float64, device=dev).index_add_(0, r, a)  # exact sums: order immaterial
    with np.errstate(divide="ignore"):
        d_host = np.power(rowsum.cpu().numpy(), -0.5)  # utils.py:210, same libm call as the reference
    d_host[np.isinf(d_host)] = 0.0                     # utils.py:211
    d = torch.from_numpy(d_host).to(dev)
    # Â[i,j] = (Ã[j,i] * d[i]) * d[j]  with Ã symmetric  (utils.py:213: adj.dot(D).transpose().dot(D))
    vals = ((a * d[r]) * d[c]).to(torch.float32)
    return r, c, vals


def _labels_and_split(n_docs: int, n_class: int, gen: torch.Generator, dev, train_frac=0.64, val_frac=0.07):
    labels = torch.randint(0, n_class, (n_docs,), generator=gen, device=dev, dtype=torch.int64)
    n_train = int(n_docs * train_frac)
    n_val = int(n_docs * val_frac)
    ar = torch.arange(n_docs, dtype=torch.int64, device=dev)
    return labels, ar[:n_train], ar[n_train:n_train + n_val], ar[n_train + n_val:]


def doc_topic_edges(n_docs: int, n_topics: int, deg_lo: int, deg_hi: int, gen: torch.Generator, dev,
                    zipf_s: float = 0.45, doc_offset: int = 0, floor: float = 0.02):
    """theta-like doc-topic edges: per document up to U{deg_lo..deg_hi} topics drawn from a Zipf(zipf_s) popularity
    (duplicates dropped; s = 0.45 spreads hub sizes ~10x like the real R8 graph, 191..1807 entries per topic row),
    weights uniform normalised to sum 1, floor 0.02 (build_graph.py:105-107)."""
    p = 1.0 / torch.arange(1, n_topics + 1, dtype=torch.float64, device=dev) ** zipf_s
    cdf = torch.cumsum(p / p.sum(), 0)
    cdf[-1] = 1.0
    slots = deg_hi
    uu = torch.rand((n_docs, slots), generator=gen, device=dev, dtype=torch.float64)
    t = torch.searchsorted(cdf, uu).clamp_(max=n_topics - 1)
    del uu
    if deg_lo < deg_hi:
        deg = torch.randint(deg_lo, deg_hi + 1, (n_docs, 1), generator=gen, device=dev)
        live = torch.arange(slots, device=dev).unsqueeze(0) < deg
    else:
        live = torch.ones((n_docs, slots), dtype=torch.bool, device=dev)
    t = torch.where(live, t, torch.full_like(t, n_topics))  # dead slots sort last
    t, _ = torch.sort(t, dim=1)
    live = t < n_topics
    live[:, 1:] &= t[:, 1:] != t[:, :-1]
    w = torch.rand((n_docs, slots), generator=gen, device=dev, dtype=torch.float32) + 0.05
    w = torch.where(live, w, torch.zeros_like(w))
    w = w / w.sum(dim=1, keepdim=True).clamp_min(1e-30)
    live &= w >= floor
    d = (torch.arange(n_docs, dtype=torch.int64, device=dev) + doc_offset).unsqueeze(1).expand(-1, slots)
    return d[live], t[live], w[live]


def topic_topic_edges(n_topics: int, gen: torch.Generator, dev, dense: bool, emb_dim: int = 100):
    """cosine similarity of random topic embeddings with a shared positive offset, keep i<j with sim > 0.3
    (build_graph.py:118-130).  `dense` -> offset large enough that (almost) every pair survives (C3/C4 shape);
    otherwise ~20% density like the real R8 graph (237 of 1225 pairs)."""
    emb = torch.randn((n_topics, emb_dim), generator=gen, device=dev, dtype=torch.float32)
    emb = emb + (1.5 if dense else 0.52)
    nrm = emb / emb.norm(dim=1, keepdim=True)
    sim = nrm @ nrm.t()
    iu = torch.triu_indices(n_topics, n_topics, offset=1, device=dev)
    s = sim[iu[0], iu[1]]
    keep = s > 0.3
    return iu[0][keep], iu[1][keep], s[keep].to(torch.float32)


def doc_topic_topic_graph(n_docs: int, n_topics: int, deg_lo: int = 8, deg_hi: int = 8, dense_topics: bool = True,
                          n_class: int = 20, seed: int = 0, device="cpu") -> SyntheticGraph:
    """C1/C2/C3/C4-shaped graph: documents 0..D-1, topics D..D+K-1."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    d, t, w = doc_topic_edges(n_docs, n_topics, deg_lo, deg_hi, gen, dev)
    ti, tj, ts = topic_topic_edges(n_topics, gen, dev, dense_topics)
    u = torch.cat([d, ti + n_docs])
    v = torch.cat([t + n_docs, tj + n_docs])
    ww = torch.cat([w, ts])
    n_dt, n_tt = int(d.numel()), int(ti.numel())
    del d, t, w, ti, tj, ts
    rows, cols, vals = normalize_undirected(u, v, ww, n_docs + n_topics)
    labels, tr, va, te = _labels_and_split(n_docs, n_class, gen, dev)
    return SyntheticGraph(n_docs, n_topics, rows, cols, vals, labels, tr, va, te, n_class,
                          meta={"doc_topic_edges": n_dt, "topic_topic_edges": n_tt, "seed": seed})


def textgcn_like_graph(n_docs: int = 7674, n_words: int = 7688, words_per_doc: int = 46,
                       word_word_pairs: int = 1_395_684, n_class: int = 8, seed: int = 0,
                       device="cpu") -> SyntheticGraph:
    """C5: TextGCN-style doc-word graph with power-law word-row skew (measured R8 figures: 323 670
    doc-word entries, 2 791 368 directed PMI>0 word-word entries, word degree median 218 / max 9 588)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    d, t, w = doc_topic_edges(n_docs, n_words, words_per_doc, words_per_doc, gen, dev, zipf_s=0.9, floor=0.0)
    w = (w * 40.0).clamp_(min=0.05)  # TF-IDF-like magnitudes
    p = 1.0 / torch.arange(1, n_words + 1, dtype=torch.float64, device=dev) ** 0.75
    cdf = torch.cumsum(p / p.sum(), 0)
    cdf[-1] = 1.0
    m = int(word_word_pairs * 1.25)
    a = torch.searchsorted(cdf, torch.rand(m, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=n_words - 1)
    b = torch.searchsorted(cdf, torch.rand(m, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=n_words - 1)
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    key = torch.unique((lo * n_words + hi)[lo != hi])[:word_word_pairs]
    wi, wj = key // n_words, key % n_words
    pmi = torch.rand(wi.numel(), generator=gen, device=dev, dtype=torch.float32) * 5.0 + 0.03
    u = torch.cat([d, wi + n_docs])
    v = torch.cat([t + n_docs, wj + n_docs])
    rows, cols, vals = normalize_undirected(u, v, torch.cat([w, pmi]), n_docs + n_words)
    labels, tr, va, te = _labels_and_split(n_docs, n_class, gen, dev)
    return SyntheticGraph(n_docs, n_words, rows, cols, vals, labels, tr, va, te, n_class,
                          meta={"doc_word_edges": int(d.numel()), "word_word_edges": int(wi.numel()), "seed": seed})


# the named BASELINE.json configurations ---------------------------------------------------------------------------
CONFIGS = {
    # name: (builder, kwargs, hidden, classes)
    "c1_r8_shape": (doc_topic_topic_graph, dict(n_docs=7674, n_topics=50, deg_lo=2, deg_hi=13, dense_topics=False, n_class=8), 200, 8),
    "c2_20ng_shape": (doc_topic_topic_graph, dict(n_docs=18846, n_topics=100, deg_lo=2, deg_hi=13, dense_topics=False, n_class=20), 200, 20),
    "c3_1m_docs_256_topics": (doc_topic_topic_graph, dict(n_docs=1_000_000, n_topics=256, deg_lo=8, deg_hi=8, dense_topics=True, n_class=20), 256, 20),
    "c4_shard_6p25m_docs_1024_topics": (doc_topic_topic_graph, dict(n_docs=6_250_000, n_topics=1024, deg_lo=8, deg_hi=8, dense_topics=True, n_class=20), 256, 20),
    "c5_textgcn_r8_shape": (textgcn_like_graph, dict(), 200, 8),
}


def make_config(name: str, device="cpu", seed: int = 0, scale: Optional[float] = None) -> tuple:
    """(graph, hidden, classes) for a named configuration; `scale` shrinks the document count (tests)."""
    builder, kw, hidden, classes = CONFIGS[name]
    kw = dict(kw)
    if scale is not None:
        kw["n_docs"] = max(64, int(kw["n_docs"] * scale))
    return builder(seed=seed, device=device, **kw), hidden, classes
