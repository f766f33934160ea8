"""Pin the CPU oracle (oracle/) against golden vectors produced by the REAL reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import gcn_oracle as O

RTOL = 2e-6  # oracle (separate mul+add) vs ATen (MKL axpy, FMA): last-bit differences only


def _adj(g):
    return O.Coo(g["adj_rows"], g["adj_cols"], g["adj_vals"], (int(g["n_docs"] + g["n_topics"]),) * 2)


def _params(g, prefix):
    return {k: g[f"{prefix}_{k}"].copy() for k in ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias")}


def _close(a, b, rtol=RTOL):
    scale = max(float(np.abs(b).max()), 1e-30)
    assert float(np.abs(a - b).max()) <= rtol * scale, (float(np.abs(a - b).max()), scale)


def test_normalize_adj_bit_exact_small(small_golden):
    g = small_golden
    n = int(g["n_docs"] + g["n_topics"])
    got = O.normalize_adj_coo(g["raw_rows"], g["raw_cols"], g["raw_vals"], n)
    assert np.array_equal(got.rows, g["adj_rows"]) and np.array_equal(got.cols, g["adj_cols"])
    assert np.array_equal(got.vals.view(np.uint32), g["adj_vals"].view(np.uint32))


def test_normalize_adj_bit_exact_r8(r8_golden):
    g = r8_golden
    n = int(g["n_docs"] + g["n_topics"])
    got = O.normalize_adj_coo(g["raw_rows"], g["raw_cols"], g["raw_vals"], n)
    assert np.array_equal(got.rows, g["adj_rows"]) and np.array_equal(got.cols, g["adj_cols"])
    assert np.array_equal(got.vals.view(np.uint32), g["adj_vals"].view(np.uint32))
    # the reference adjacency is exactly symmetric
    rp, ci, v = O.csr_from_coo(got)
    rp2, ci2, v2 = O.csr_from_coo(got.transpose())
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and np.array_equal(v.view(np.uint32), v2.view(np.uint32))


def test_spmm_matches_reference(small_golden):
    g = small_golden
    adj = _adj(g)
    _close(O.spmm(adj, g["spmm_B"]), g["spmm_Y"])
    rp, ci, v = O.csr_from_coo(adj)
    y_csr = O.spmm_csr(rp, ci, v, g["spmm_B"], n_threads=2)
    assert np.array_equal(y_csr.view(np.uint32), O.spmm(adj, g["spmm_B"]).view(np.uint32))
    _close(O.spmm_f64(adj, g["spmm_B"]).astype(np.float32), g["spmm_Y"], rtol=1e-6)


def test_csr_from_coo_coalesce_semantics():
    rng = np.random.default_rng(0)
    n, m = 37, 41
    rows = rng.integers(0, n, 600)
    cols = rng.integers(0, m, 600)
    vals = rng.normal(size=600).astype(np.float32)
    rp, ci, v = O.csr_from_coo(O.Coo(rows, cols, vals, (n, m)))
    import torch
    t = torch.sparse_coo_tensor(torch.tensor(np.stack([rows, cols])), torch.tensor(vals), (n, m)).coalesce()
    assert np.array_equal(ci, t.indices()[1].numpy().astype(np.int32))
    counts = np.bincount(t.indices()[0].numpy(), minlength=n)
    assert np.array_equal(np.diff(rp), counts)
    np.testing.assert_allclose(v, t.values().numpy(), rtol=1e-6, atol=1e-7)


def test_graph_convolution_layer(small_golden):
    g = small_golden
    n = int(g["n_docs"] + g["n_topics"])
    x = O.Coo(g["sf_x_rows"], g["sf_x_cols"], g["sf_x_vals"], (n, 24))
    _close(O.graph_convolution(x, _adj(g), g["gc_weight"], g["gc_bias"]), g["gc_out"])


@pytest.mark.parametrize("prefix", ["fl", "sf"])
def test_gcn_forward_backward(small_golden, prefix):
    g = small_golden
    n = int(g["n_docs"] + g["n_topics"])
    adj = _adj(g)
    x = None if prefix == "fl" else O.Coo(g["sf_x_rows"], g["sf_x_cols"], g["sf_x_vals"], (n, 24))
    params = _params(g, prefix)
    logits, _ = O.gcn_forward(x, adj, params, p=float(g["p"]), training=False)
    _close(logits, g[f"{prefix}_eval_logits"])
    # train mode with the mask torch's th.dropout drew (proves the mask reconstruction used by the paired tests)
    loss, logits, grads = O.gcn_loss_and_grads(x, adj, params, g["target"], g["index"], p=float(g["p"]),
                                               training=True, keep_mask=g[f"{prefix}_keep_mask"])
    _close(logits, g[f"{prefix}_train_logits"])
    assert abs(loss - float(g[f"{prefix}_train_loss"])) <= 2e-6 * max(1.0, abs(loss))
    for k in ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"):
        _close(grads[k], g[f"{prefix}_grad_{k}"], rtol=5e-6)


def test_r8_seed0_step(r8_golden):
    """Real R8 TopicGCN graph, featureless, seed 0: reference logits / loss / gradients."""
    import torch
    from tests.golden.make_golden_shared import mask_seed, reference_init, train_val_split
    g = r8_golden
    n = int(g["n_docs"] + g["n_topics"])
    adj = _adj(g)
    params = reference_init(0, n, int(g["nhid"]), int(g["nclass"]))
    train_lst, _ = train_val_split(g["train_all"], 0)
    logits, _ = O.gcn_forward(None, adj, params, training=False)
    _close(logits, g["s0_eval_logits"], rtol=5e-6)
    torch.manual_seed(mask_seed(0, 0))
    mask = torch.empty(n, int(g["nhid"])).bernoulli_(0.5).numpy().astype(np.uint8)
    loss, logits, grads = O.gcn_loss_and_grads(None, adj, params, g["target"].astype(np.int64), train_lst, p=0.5,
                                               training=True, keep_mask=mask)
    assert abs(loss - float(g["s0_train_loss"])) <= 5e-6 * max(1.0, abs(loss))
    _close(logits[::97], g["s0_train_logits_rows"], rtol=5e-6)
    _close(grads["gc2.weight"], g["s0_grad_gc2.weight"], rtol=2e-5)
    _close(grads["gc1.bias"], g["s0_grad_gc1.bias"], rtol=2e-5)
    _close(grads["gc1.weight"][int(g["n_docs"]):], g["s0_grad_gc1.weight_topics"], rtol=2e-5)


def test_philox_mask_statistics():
    m = O.philox_keep_mask(257, 200, 0.5, seed=1234, offset=3)
    assert m.shape == (257, 200) and abs(m.mean() - 0.5) < 0.01
    m2 = O.philox_keep_mask(257, 200, 0.5, seed=1234, offset=4)
    assert (m != m2).mean() > 0.4
    assert abs(O.philox_keep_mask(64, 256, 0.2, 7, 1).mean() - 0.8) < 0.02


def test_metrics_match_reference_golden():
    """oracle class_counts / metrics_from_counts against utils.accuracy / utils.macro_f1 of the real reference
    (tests/golden/make_golden_metrics.py), including classes that never occur (0/0 -> 0)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.npz"))
    for i in range(int(g["n_cases"])):
        logits, target, ref = g[f"logits_{i}"], g[f"target_{i}"], g[f"ref_{i}"]
        index = np.arange(logits.shape[0])
        counts = O.class_counts(logits, target, index, logits.shape[1])
        m = O.metrics_from_counts(counts, index.size)
        got = np.array([m["acc"], m["macro_f1"], m["precision"], m["recall"]])
        assert np.allclose(got, ref, rtol=0, atol=1e-12), (i, got, ref)


@pytest.mark.parametrize("prefix", ["fl", "sf"])
def test_torch_ref_matches_reference_golden(small_golden, prefix):
    """oracle/torch_ref.py (the reference's module restated on the same ATen operators: the ATen CPU baseline and the
    cuSPARSE GPU baseline of bench.py) against vectors the REAL reference modules produced: eval logits and loss, and in
    train mode with dropout off (p = 0 draws no mask) the gradients of the oracle's own backward."""
    import torch
    from oracle import torch_ref as TR
    g = small_golden
    n = int(g["n_docs"] + g["n_topics"])
    adj = torch.sparse_coo_tensor(torch.tensor(np.stack([g["adj_rows"], g["adj_cols"]])), torch.tensor(g["adj_vals"]), (n, n))
    if prefix == "fl":
        x, nfeat = TR.sparse_identity(n, "cpu"), n
    else:
        x = torch.sparse_coo_tensor(torch.tensor(np.stack([g["sf_x_rows"], g["sf_x_cols"]])), torch.tensor(g["sf_x_vals"]), (n, 24))
        nfeat = 24
    model = TR.GCNRef(nfeat, int(g["nhid"]), int(g["nclass"]), 0.0)
    model.load_state_dict({k: torch.tensor(g[f"{prefix}_{k}"]) for k in ("gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias")})
    model.eval()
    with torch.no_grad():
        _close(model(x, adj).numpy(), g[f"{prefix}_eval_logits"])
    model.train()
    target, index = torch.tensor(g["target"]), torch.tensor(g["index"])
    loss = TR.train_step(model, x, adj, target, index)
    params = _params(g, prefix)
    xo = None if prefix == "fl" else O.Coo(g["sf_x_rows"], g["sf_x_cols"], g["sf_x_vals"], (n, 24))
    ref_loss, _, ref_grads = O.gcn_loss_and_grads(xo, _adj(g), params, g["target"], g["index"], p=0.0, training=True, keep_mask=None)
    assert abs(float(loss) - float(ref_loss)) <= 2e-6 * max(1.0, abs(float(ref_loss)))
    for k, p in model.named_parameters():
        _close(p.grad.numpy(), ref_grads[k], rtol=5e-6)
