"""Document-row sharding of the doc-topic-topic graph over G GPUs (one process per GPU, NCCL over NVLink).

Partition.  Documents are row-sharded contiguously: rank g owns D_g documents together with their rows
of Â, W1 (featureless), H1 and their labels.  The K topic nodes are REPLICATED: every rank keeps the K topic rows of
every dense operand, and a local square adjacency over its (D_g + K) nodes:

    document rows : complete        (self loop + topic columns: everything a document row touches is local)
    topic rows    : local doc columns only, plus this rank's column slice of the K x K topic-topic block
                    (the slices tile the block, so summing the ranks' partial topic rows gives the full rows)

so  Y = Â B  on rank g is the local SpMM followed by ONE all-reduce (sum) of the K x F topic rows of Y — the only
collective of a layer.  Â is symmetric, hence the backward products Â^T dZ are the same operation.  Replicated
parameters (W2, b1, b2) get one packed all-reduce of their gradients (+ the loss) per step; W1 is row-sharded with
the documents and its K topic rows stay bit-identical on all ranks because they only ever see all-reduced values.
Per train step: three all-reduces (K x H forward, K x C backward, one packed [K x H | dW2 | db1 | db2 | loss] at the
end), the first two issued on a side stream behind document-row kernels that do not depend on them.

The local SpMM runs the fused kernels of the single-GPU path; rows >= D_g (the topic rows) are stored raw
(`raw_row_begin`), all-reduced, and then given their epilogue.  The reference has nothing comparable (single process,
single device: trainer.py:425); the numerical contract is that the sharded result equals the single-GPU result up to
the summation grouping of the topic rows.

`Comm` abstracts the collective: `TorchDistComm` (torch.distributed, NCCL on GPUs / gloo in the CPU tests) and
`ThreadComm` (ranks emulated as threads of one process sharing one device — used to test the sharded step on a
single GPU without waiting kernels).
"""
from __future__ import annotations

import functools
import threading
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch

from . import graphgen


# ---------------------------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------------------------
class TorchDistComm:
    """torch.distributed process group (NCCL or gloo)."""

    overlap = True  # collectives may be issued on a side stream (they order themselves against the stream they are called on)

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        self.dist.broadcast(t, src=src, group=self.group)
        return t


class SingleComm:
    rank, world = 0, 1

    def all_reduce(self, t):
        return t

    def broadcast(self, t, src=0):
        return t


class ThreadComm:
    """Ranks emulated as threads of ONE process on one device: an all-reduce is 'everybody deposits, barrier, everybody
    adds the deposits in rank order, barrier'.  All work is enqueued on the same CUDA stream, so stream order makes the
    deposits visible to the sums; no kernel ever waits on another."""

    class _Shared:
        def __init__(self, world):
            self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)

    def __init__(self, shared: "_Shared", rank: int):
        self.s, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def create(cls, world: int) -> List["ThreadComm"]:
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        self.s.slots[self.rank] = t.clone()
        self.s.barrier.wait()
        acc = self.s.slots[0].clone()
        for r in range(1, self.world):
            acc += self.s.slots[r]
        self.s.barrier.wait()
        t.copy_(acc)
        return t

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.rank == src:
            self.s.slots[src] = t.clone()
        self.s.barrier.wait()
        t.copy_(self.s.slots[src])
        self.s.barrier.wait()
        return t


# ---------------------------------------------------------------------------------------------------------------
# local graph of one rank
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class LocalGraph:
    """Row-major COO of the rank-local square adjacency over (local documents | all topics)."""
    rank: int
    world: int
    n_docs_local: int
    n_topics: int
    rows: torch.Tensor
    cols: torch.Tensor
    vals: torch.Tensor
    n_docs_global: int = 0
    nnz_global: int = 0
    labels: Optional[torch.Tensor] = None
    train_idx: Optional[torch.Tensor] = None
    n_train_global: int = 0
    n_class: int = 0
    csr: object = None  # DeviceCSR, built lazily on CUDA

    @property
    def n_local(self) -> int:
        return self.n_docs_local + self.n_topics

    @property
    def nnz(self) -> int:
        return int(self.rows.numel())


def topic_slice(n_topics: int, rank: int, world: int):
    """Column slice [lo, hi) of the topic-topic block that rank owns."""
    return (rank * n_topics) // world, ((rank + 1) * n_topics) // world


def build_local_graph(doc: torch.Tensor, topic: torch.Tensor, w: torch.Tensor, tt_i: torch.Tensor, tt_j: torch.Tensor,
                      tt_s: torch.Tensor, n_docs_local: int, n_topics: int, comm) -> LocalGraph:
    """Normalise (reference utils.py:206-213 arithmetic, float64) and lay out one rank's local adjacency.

    doc/topic/w : this rank's document-topic edges (local document index, topic index, fp32 weight)
    tt_*        : ALL topic-topic edges i<j with fp32 similarity (identical on every rank)
    The topic degrees need every rank's documents: one float64 all-reduce of K sums (exact in float64, so the result
    does not depend on the reduction order)."""
    dev = doc.device
    D, K = int(n_docs_local), int(n_topics)
    w64 = w.to(torch.float64)
    deg_doc = torch.ones(D, dtype=torch.float64, device=dev).index_add_(0, doc, w64)
    deg_top = torch.zeros(K, dtype=torch.float64, device=dev).index_add_(0, topic, w64)
    comm.all_reduce(deg_top)
    s64 = tt_s.to(torch.float64)
    deg_top = deg_top + 1.0
    deg_top.index_add_(0, tt_i, s64).index_add_(0, tt_j, s64)
    with np.errstate(divide="ignore"):
        dd = np.power(deg_doc.cpu().numpy(), -0.5)  # utils.py:210 (same libm call as the reference)
        dt = np.power(deg_top.cpu().numpy(), -0.5)
    dd[np.isinf(dd)] = 0.0
    dt[np.isinf(dt)] = 0.0
    dd, dt = torch.from_numpy(dd).to(dev), torch.from_numpy(dt).to(dev)
    lo, hi = topic_slice(K, comm.rank, comm.world)
    ar_d = torch.arange(D, dtype=torch.int64, device=dev)
    # Â[i,j] = (Ã[j,i] * d_i) * d_j
    r = [ar_d, doc, topic + D]
    c = [ar_d, topic + D, doc]
    v = [(dd * dd), (w64 * dd[doc]) * dt[topic], (w64 * dt[topic]) * dd[doc]]
    # this rank's column slice of the topic-topic block (both directions) and of the topic self loops
    for a, b in ((tt_i, tt_j), (tt_j, tt_i)):
        keep = (b >= lo) & (b < hi)
        r.append(a[keep] + D)
        c.append(b[keep] + D)
        v.append((s64[keep] * dt[a[keep]]) * dt[b[keep]])
    ks = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    r.append(ks + D)
    c.append(ks + D)
    v.append(dt[ks] * dt[ks])
    rows, cols, vals = torch.cat(r), torch.cat(c), torch.cat(v).to(torch.float32)
    order = torch.argsort(rows * (D + K) + cols)
    return LocalGraph(comm.rank, comm.world, D, K, rows[order], cols[order], vals[order])


def make_sharded_config(name: str, rank: int, world: int, device, seed: int = 0, comm=None,
                        docs_per_rank: Optional[int] = None) -> LocalGraph:
    """One rank's shard of a named synthetic configuration: `docs_per_rank` documents (default: the configuration's
    document count, i.e. weak scaling) generated from a rank-specific seed, topics and topic-topic edges from the shared
    seed."""
    builder, kw, hidden, n_class = graphgen.CONFIGS[name]
    if builder is not graphgen.doc_topic_topic_graph:
        raise ValueError(f"{name} is not a document-topic-topic configuration")
    kw = dict(kw)
    D = int(docs_per_rank or kw["n_docs"])
    K = int(kw["n_topics"])
    dev = torch.device(device)
    comm = comm or (TorchDistComm() if world > 1 else SingleComm())
    g_shared = torch.Generator(device=dev).manual_seed(seed)
    tt_i, tt_j, tt_s = graphgen.topic_topic_edges(K, g_shared, dev, kw["dense_topics"])
    g_rank = torch.Generator(device=dev).manual_seed(seed * 1000 + 17 + rank)
    d, t, w = graphgen.doc_topic_edges(D, K, kw["deg_lo"], kw["deg_hi"], g_rank, dev)
    lg = build_local_graph(d, t, w, tt_i, tt_j, tt_s, D, K, comm)
    labels, tr, _va, _te = graphgen._labels_and_split(D, n_class, g_rank, dev)
    lg.labels, lg.train_idx, lg.n_class = labels, tr, n_class
    cnt = torch.tensor([float(D), float(tr.numel()), float(2 * d.numel() + D)], dtype=torch.float64, device=dev)
    comm.all_reduce(cnt)
    lg.n_docs_global, lg.n_train_global = int(cnt[0].item()), int(cnt[1].item())
    lg.nnz_global = int(cnt[2].item()) + 2 * int(tt_i.numel()) + K
    return lg


def shard_edges(n_docs: int, world: int):
    """Contiguous document ranges [lo, hi) per rank."""
    return [((r * n_docs) // world, ((r + 1) * n_docs) // world) for r in range(world)]


# ---------------------------------------------------------------------------------------------------------------
# the sharded SpMM and the sharded train step (CUDA)
# ---------------------------------------------------------------------------------------------------------------
def _csr(lg: LocalGraph):
    if lg.csr is None:
        from .csr import DeviceCSR
        lg.csr = DeviceCSR.from_coo(lg.rows, lg.cols, lg.vals, lg.n_local, lg.n_local, symmetric=True)
    return lg.csr


def sharded_spmm(lg: LocalGraph, B: torch.Tensor, comm, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = Â B over the sharded graph: local SpMM + all-reduce of the K topic rows."""
    from . import ops
    Y = ops.spmm(_csr(lg), B, None, out=out)
    comm.all_reduce(Y[lg.n_docs_local:])
    return Y


class ShardedGCN(torch.nn.Module):
    """Featureless 2-layer GCN on one rank's shard.  Parameters: gc1.weight [D_g + K, nhid] (document rows are this
    rank's, topic rows replicated), gc1.bias, gc2.weight, gc2.bias (replicated).  Same init distribution as the
    reference layers (layer.py:67-82); replicated tensors are broadcast from rank 0 so all ranks start identical."""

    def __init__(self, lg: LocalGraph, nhid: int, nclass: int, dropout: float, comm=None):
        super().__init__()
        from .layer import GraphConvolution
        self.lg = lg
        self.comm = comm or (TorchDistComm() if lg.world > 1 else SingleComm())
        self.gc1 = GraphConvolution(lg.n_local, nhid)
        self.gc2 = GraphConvolution(nhid, nclass)
        self.dropout = float(dropout)
        self._seed: Optional[int] = None
        self._calls = 0
        self._side: Optional[torch.cuda.Stream] = None

    def sync_replicated(self) -> None:
        """Make the replicated tensors identical on all ranks (call after moving the module to its device)."""
        with torch.no_grad():
            D = self.lg.n_docs_local
            top = self.gc1.weight.data[D:].contiguous()
            self.comm.broadcast(top)
            self.gc1.weight.data[D:] = top
            for p in (self.gc1.bias, self.gc2.weight, self.gc2.bias):
                self.comm.broadcast(p.data)

    def set_dropout_seed(self, seed: int) -> None:
        """The shared dropout seed (must be the same on every rank)."""
        self._seed, self._calls = int(seed), 0

    def dropout_seeds(self):
        """(document-row seed of this rank, topic-row seed shared by all ranks).  The shared seed is drawn from torch's
        global generator on rank 0 (so `th.manual_seed` makes runs reproducible, like GCN._dropout_state) and broadcast;
        the Philox counter is keyed on the LOCAL row index, so each rank mixes its rank into the seed of its document
        rows — otherwise document i of every shard would draw the same mask.  The replicated topic rows keep the shared
        seed: their masks must agree on all ranks."""
        if self._seed is None:
            t = torch.randint(0, 2**62, (1,), dtype=torch.int64).to(self.gc1.weight.device)
            self.comm.broadcast(t)
            self._seed = int(t.item())
        rank = int(getattr(self.comm, "rank", 0))
        return (self._seed ^ ((rank + 1) * 0x9E3779B97F4A7C15)) & (2**63 - 1), self._seed ^ 0x7091C

    def side_stream(self) -> Optional[torch.cuda.Stream]:
        """Stream for the collectives and the topic-row work that overlap the document-row kernels; None when the
        communicator emulates ranks on one stream (ThreadComm) or there is nothing to overlap (one rank)."""
        if not getattr(self.comm, "overlap", False):
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(self.gc1.weight.device)
        return self._side

    def loss(self, labels: Optional[torch.Tensor] = None, index: Optional[torch.Tensor] = None,
             row_label: Optional[torch.Tensor] = None, keep_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Global mean cross-entropy over all ranks' training documents (reference trainer.py:357-359 semantics).

        In training mode with gradients enabled the sum over ranks of the loss travels in the LAST all-reduce of the
        backward pass (one collective fewer per step): the returned tensor holds this rank's partial sum until
        `backward()` has run, and the global value from then on — the order the reference trainer uses it in
        (`loss.backward(); ...; loss.item()`, trainer.py:360-367).  Without gradients the loss is reduced at once."""
        from . import ops
        lg = self.lg
        if row_label is None:
            labels = lg.labels if labels is None else labels
            index = lg.train_idx if index is None else index
            if labels.is_cuda and index.is_cuda:
                row_label = ops.make_row_label(lg.n_local, labels, index)
            else:  # host labels: a callable, so that the copies are queued (side stream) only after layer 1 has been launched
                row_label = functools.partial(ops.make_row_label_async, lg.n_local, labels, index, self.gc1.weight.device)
        self._calls += int(self.training)
        defer = bool(self.training and torch.is_grad_enabled() and self.comm.world > 1)  # (decided here: grad mode is off inside Function.forward)
        return _ShardedLoss.apply(self.gc1.weight, self.gc1.bias, self.gc2.weight, self.gc2.bias, self, row_label,
                                  keep_mask, defer)


class _Fork:
    """`with _Fork(side):` runs its body on the side stream after everything queued on the current stream so far; the
    caller joins with `.join()` (current stream waits for the side stream).  side = None: the body runs in line."""

    def __init__(self, side: Optional[torch.cuda.Stream]):
        self.side, self.ctx = side, None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)


def sharded_forward(model: ShardedGCN, W1, b1, W2, b2, row_label, keep_mask, defer_loss: bool = False):
    """Train-mode forward on one rank.  Returns (loss, saved tensors for the backward); the loss is the global mean
    unless `defer_loss` (then this rank's partial sum: sharded_backward adds the ranks up in its last collective).

    Collectives: ONE all-reduce of the K x H topic rows, issued on the side stream together with the topic-row epilogue and
    the topic rows of S2 = H1 W2, while the main stream computes the document rows of S2 (0.25 ms at 1 M documents: the
    latency-bound all-reduce disappears behind it)."""
    from . import ops
    lg, comm = model.lg, model.comm
    csr, D, K = _csr(lg), lg.n_docs_local, lg.n_topics
    p, training = model.dropout, model.training
    inv = 1.0 / max(lg.n_train_global, 1)
    seed_doc, seed_top = model.dropout_seeds() if (training and p > 0 and keep_mask is None) else (0, 0)
    # layer 1: document rows get the fused epilogue, topic rows are stored raw, summed over ranks, then finished
    H1 = ops.gc1_forward(csr, W1, b1, p, training, keep_mask, seed_doc, model._calls, raw_row_begin=D)
    if callable(row_label):  # host labels: start their copy only now, with layer 1 already queued on the device
        row_label = row_label()
    S2 = torch.empty((lg.n_local, int(W2.shape[1])), dtype=torch.float32, device=H1.device)
    top = H1[D:]
    with _Fork(model.side_stream()) as side:
        comm.all_reduce(top)
        ops.gc1_forward(ops.identity_csr(K, H1.device), top, b1, p, training,
                        None if keep_mask is None else keep_mask[D:].contiguous(), seed_top, model._calls, out=top)
        ops.dense_nn(top, W2, out=S2[D:])
    ops.dense_nn(H1[:D], W2, out=S2[:D])
    side.join()
    # layer 2 + loss: only document rows carry labels, so their logits need no collective in the forward
    loss, _, dZ2 = ops.gc2_loss_forward(csr, S2, b2, row_label, inv, want_logits=False, want_grad=True)
    if not defer_loss:
        loss = loss.clone()
        comm.all_reduce(loss)
    return loss, (H1, dZ2)


def sharded_backward(model: ShardedGCN, W2, saved, dloss=None, loss_partial: Optional[torch.Tensor] = None):
    """Backward on one rank: returns (dW1, db1, dW2, db2), replicated gradients already summed over ranks.

    Collectives: the K x C topic rows of dS2 (side stream, behind the column sums that give db2) and ONE
    packed all-reduce at the end: the K x H topic rows of dW1, dW2, db1, db2 and — when `loss_partial` is given — the
    deferred loss sum, which is written back into `loss_partial` in place.  `dloss` (the upstream gradient of the
    scalar loss, a device scalar) is folded into dS2 by the SpMM epilogue and into db2 on C elements: no pass over dZ2."""
    from . import ops
    lg, comm = model.lg, model.comm
    csr, D, K = _csr(lg), lg.n_docs_local, lg.n_topics
    H1, dZ2 = saved
    g = None if dloss is None else dloss.reshape(()).to(torch.float32).contiguous()
    scale = 1.0 / (1.0 - model.dropout) if (model.training and model.dropout > 0) else 1.0
    dS2 = ops.spmm(csr, dZ2, out_scale=g)
    with _Fork(model.side_stream()) as side:
        comm.all_reduce(dS2[D:])
    db2 = ops.colsum(dZ2)  # (independent of the all-reduce: hides its latency; topic rows of dZ2 are zero)
    if g is not None:
        db2 = db2 * g
    side.join()
    # one launch over all rows; the replicated topic rows give their dZ1 everywhere but count into dW2 / db1 on rank 0 only
    dZ1, dW2, db1 = ops.hidden_backward(H1, dS2, W2, scale, n_count=(lg.n_local if comm.rank == 0 else D))
    dW1 = ops.spmm(csr, dZ1)
    h, c = dW2.shape
    parts = [dW1[D:].reshape(-1), dW2.reshape(-1), db1, db2]
    if loss_partial is not None:
        parts.append(loss_partial.reshape(1))
    packed = torch.cat(parts)
    comm.all_reduce(packed)
    o = K * h
    dW1[D:] = packed[:o].view(K, h)
    dW2, db1, db2 = packed[o:o + h * c].view(h, c), packed[o + h * c:o + h * c + h], packed[o + h * c + h:o + h * c + h + c]
    if loss_partial is not None:
        loss_partial.copy_(packed[-1])
    return dW1, db1, dW2, db2


class _ShardedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W1, b1, W2, b2, model, row_label, keep_mask, defer):
        loss, saved = sharded_forward(model, W1, b1, W2, b2, row_label, keep_mask, defer_loss=defer)
        ctx.model = model
        ctx.loss_partial = loss.detach() if defer else None  # (an alias of the output's storage, without its autograd node)
        ctx.save_for_backward(W2, *saved)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        W2, H1, dZ2 = ctx.saved_tensors
        dW1, db1, dW2, db2 = sharded_backward(ctx.model, W2, (H1, dZ2), dloss, loss_partial=ctx.loss_partial)
        return dW1, db1, dW2, db2, None, None, None, None


def run_threads(world: int, fn: Callable[[int, ThreadComm], object]) -> list:
    """Run fn(rank, comm) for `world` emulated ranks as threads; returns their results (exceptions re-raised)."""
    comms = ThreadComm.create(world)
    out: list = [None] * world
    err: list = [None] * world

    def target(r):
        try:
            out[r] = fn(r, comms[r])
        except BaseException as exc:  # noqa: BLE001
            err[r] = exc
            try:
                comms[r].s.barrier.abort()
            except Exception:
                pass

    ts = [threading.Thread(target=target, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out
