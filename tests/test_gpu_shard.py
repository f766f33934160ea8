"""Document-sharded train step on the GPU: G emulated ranks (threads sharing one device, ThreadComm) and, when the box
has >= 2 GPUs, real NCCL ranks, must reproduce the single-GPU step (loss and all four gradients)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D, K, H, C = 3000, 24, 64, 8


def _edges(dev):
    from topicgcn_b200 import graphgen
    gen = torch.Generator(device=dev).manual_seed(5)
    d, t, w = graphgen.doc_topic_edges(D, K, 2, 9, gen, dev)
    ti, tj, ts = graphgen.topic_topic_edges(K, gen, dev, dense=True)
    return d, t, w, ti, tj, ts


def _reference_step(dev, params, labels, train_idx, mask):
    """Single-GPU step through the public module on the global graph."""
    import topicgcn_b200 as tg
    from topicgcn_b200 import graphgen
    d, t, w, ti, tj, ts = _edges(dev)
    u = torch.cat([d, ti + D]); v = torch.cat([t + D, tj + D]); ww = torch.cat([w, ts])
    r, c, vals = graphgen.normalize_undirected(u, v, ww, D + K)
    adj = torch.sparse_coo_tensor(torch.stack([r, c]), vals, (D + K, D + K), check_invariants=False)
    model = tg.GCN(D + K, H, C, 0.5).to(dev)
    model.load_state_dict(params)
    model.train()
    model.set_next_dropout_mask(mask)
    loss = model.loss(tg.Featureless(D + K), adj, labels, train_idx)
    loss.backward()
    return float(loss), {k: p.grad.clone() for k, p in model.named_parameters()}


def _rank_step(rank, comm, world, dev, params, labels, train_idx, mask):
    from topicgcn_b200 import ops, shard
    d, t, w, ti, tj, ts = _edges(dev)
    lo, hi = shard.shard_edges(D, world)[rank]
    mine = (d >= lo) & (d < hi)
    lg = shard.build_local_graph(d[mine] - lo, t[mine], w[mine], ti, tj, ts, hi - lo, K, comm)
    lg.n_train_global = int(train_idx.numel())
    model = shard.ShardedGCN(lg, H, C, 0.5, comm=comm).to(dev)
    with torch.no_grad():
        model.gc1.weight.copy_(torch.cat([params["gc1.weight"][lo:hi], params["gc1.weight"][D:]]))
        model.gc1.bias.copy_(params["gc1.bias"]); model.gc2.weight.copy_(params["gc2.weight"]); model.gc2.bias.copy_(params["gc2.bias"])
    model.train()
    tr_local = train_idx[(train_idx >= lo) & (train_idx < hi)] - lo
    row_label = ops.make_row_label(lg.n_local, labels[lo:hi], tr_local)
    mask_local = torch.cat([mask[lo:hi], mask[D:]]).contiguous()
    loss, saved = shard.sharded_forward(model, model.gc1.weight.data, model.gc1.bias.data, model.gc2.weight.data,
                                        model.gc2.bias.data, row_label, mask_local)
    dW1, db1, dW2, db2 = shard.sharded_backward(model, model.gc2.weight.data, saved)
    return lo, hi, float(loss), dW1.clone(), db1.clone(), dW2.clone(), db2.clone()


def _check(ref_loss, ref_g, results):
    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    for lo, hi, loss, dW1, db1, dW2, db2 in results:
        assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
        Dl = hi - lo
        assert rel(dW1[:Dl], ref_g["gc1.weight"][lo:hi]) <= 2e-5     # this rank's document rows
        assert rel(dW1[Dl:], ref_g["gc1.weight"][D:]) <= 2e-5        # replicated topic rows
        assert rel(db1, ref_g["gc1.bias"]) <= 2e-5 and rel(dW2, ref_g["gc2.weight"]) <= 2e-5
        assert rel(db2, ref_g["gc2.bias"]) <= 2e-5
    # replicated results are bit-identical on every rank
    for other in results[1:]:
        assert torch.equal(other[4], results[0][4]) and torch.equal(other[5], results[0][5])
        assert torch.equal(other[3][other[1] - other[0]:], results[0][3][results[0][1] - results[0][0]:])


def _inputs(dev):
    import topicgcn_b200 as tg
    torch.manual_seed(3)
    params = {k: v.to(dev) for k, v in tg.GCN(D + K, H, C, 0.5).state_dict().items()}
    g = torch.Generator(device="cpu").manual_seed(9)
    labels = torch.randint(0, C, (D,), generator=g).to(dev)
    train_idx = torch.sort(torch.randperm(D, generator=g)[: int(0.6 * D)]).values.to(dev)
    mask = torch.empty(D + K, H).bernoulli_(0.5, generator=g).to(torch.uint8).to(dev)
    return params, labels, train_idx, mask


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_step_matches_single_gpu_emulated_ranks(world):
    from topicgcn_b200 import shard
    dev = torch.device("cuda:0")
    params, labels, train_idx, mask = _inputs(dev)
    ref_loss, ref_g = _reference_step(dev, params, labels, train_idx, mask)
    results = shard.run_threads(world, lambda r, comm: _rank_step(r, comm, world, dev, params, labels, train_idx, mask))
    _check(ref_loss, ref_g, results)


def _nccl_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from topicgcn_b200 import shard
    params, labels, train_idx, mask = _inputs(dev)
    out = _rank_step(rank, shard.TorchDistComm(), world, dev, params, labels, train_idx, mask)
    ret[rank] = tuple(o.cpu() if torch.is_tensor(o) else o for o in out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_step_matches_single_gpu_nccl():
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(world, port, ret), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    params, labels, train_idx, mask = _inputs(dev)
    ref_loss, ref_g = _reference_step(dev, params, labels, train_idx, mask)
    results = [tuple(o.to(dev) if torch.is_tensor(o) else o for o in ret[r]) for r in range(world)]
    _check(ref_loss, ref_g, results)
