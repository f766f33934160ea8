"""N>1 host logic on CPU: two gloo ranks build their local shards (one float64 all-reduce of the topic degrees),
multiply with the oracle, all-reduce the K topic rows, and must reproduce the global product."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D, K, F = 600, 12, 16


def _edges():
    from topicgcn_b200 import graphgen
    gen = torch.Generator().manual_seed(11)
    dev = torch.device("cpu")
    d, t, w = graphgen.doc_topic_edges(D, K, 2, 9, gen, dev)
    ti, tj, ts = graphgen.topic_topic_edges(K, gen, dev, dense=True)
    return d, t, w, ti, tj, ts


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gcn_oracle as O
    from topicgcn_b200 import shard
    d, t, w, ti, tj, ts = _edges()
    lo, hi = shard.shard_edges(D, world)[rank]
    mine = (d >= lo) & (d < hi)
    comm = shard.TorchDistComm()
    lg = shard.build_local_graph(d[mine] - lo, t[mine], w[mine], ti, tj, ts, hi - lo, K, comm)
    B = torch.tensor(np.random.default_rng(0).normal(size=(D + K, F)).astype(np.float32))
    B_loc = torch.cat([B[lo:hi], B[D:]])
    coo = O.Coo(lg.rows.numpy(), lg.cols.numpy(), lg.vals.numpy(), (lg.n_local, lg.n_local))
    Y = torch.tensor(O.spmm(coo, B_loc.numpy()))
    comm.all_reduce(Y[lg.n_docs_local:])  # the one collective of a layer
    ret[rank] = (lo, hi, Y.numpy(), lg.rows.numpy(), lg.cols.numpy(), lg.vals.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_spmm_matches_global():
    from oracle import gcn_oracle as O
    from topicgcn_b200 import graphgen
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    d, t, w, ti, tj, ts = _edges()
    u = torch.cat([d, ti + D]); v = torch.cat([t + D, tj + D]); ww = torch.cat([w, ts])
    r, c, vals = graphgen.normalize_undirected(u, v, ww, D + K)
    glob = O.Coo(r.numpy(), c.numpy(), vals.numpy(), (D + K, D + K))
    B = np.random.default_rng(0).normal(size=(D + K, F)).astype(np.float32)
    Yg = O.spmm(glob, B)
    nnz_local = 0
    for rank in range(world):
        lo, hi, Y, lr, lc, lv = ret[rank]
        Dl = hi - lo
        # document rows are exact, topic rows equal up to the grouping of the cross-rank sum
        assert np.array_equal(Y[:Dl], Yg[lo:hi])
        assert np.abs(Y[Dl:] - Yg[D:]).max() <= 1e-5 * np.abs(Yg[D:]).max()
        # the local matrices tile the global one: every local entry is a global entry with the same bits
        gr = np.where(lr < Dl, lr + lo, lr - Dl + D)
        gc = np.where(lc < Dl, lc + lo, lc - Dl + D)
        key_g = dict(zip((r.numpy() * (D + K) + c.numpy()).tolist(), vals.numpy().view(np.uint32).tolist()))
        for k_, v_ in zip((gr * (D + K) + gc).tolist(), lv.view(np.uint32).tolist()):
            assert key_g[k_] == v_
        nnz_local += lr.size
    assert nnz_local == r.numel()  # ... and every global entry is owned by exactly one rank
