"""Golden fixture for the edge-list ingest, produced by the REAL reference code path in the build container:
    python tests/golden/make_golden_ingest.py
Writes tests/golden/edges_small.txt (the `u v w` text format of nx.write_weighted_edgelist, build_graph.py:199 — with a
repeated edge, a reversed duplicate and shuffled lines appended by hand) and tests/golden/edges_small_adj.npz: the
normalised adjacency the reference's ingest makes of that file (trainer.py:98-151: nx.read_weighted_edgelist(nodetype=int)
-> nx.adjacency_matrix(nodelist=range(n), dtype=float32) -> symmetrise -> utils.preprocess_adj)."""
import os
import sys
import types

import networkx as nx
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)
if not hasattr(np, "Inf"):
    np.Inf = np.inf
if "prettytable" not in sys.modules:
    stub = types.ModuleType("prettytable")
    stub.PrettyTable = object
    sys.modules["prettytable"] = stub
import utils as ref_utils  # noqa: E402  (the reference module)
from make_golden import small_graph  # noqa: E402


def main():
    g, n_docs, n_topics = small_graph(seed=5, n_docs=400, n_topics=16)
    path = os.path.join(HERE, "edges_small.txt")
    nx.write_weighted_edgelist(g, path)  # build_graph.py:199
    with open(path) as fh:
        lines = fh.read().splitlines()
    rng = np.random.default_rng(1)
    rng.shuffle(lines)
    u, v, w = lines[3].split()
    lines.append(f"{u} {v} 0.4375")        # the same edge again: the last weight wins (nx.Graph.add_edge overwrites)
    u2, v2, w2 = lines[10].split()
    lines.append(f"{v2} {u2} 0.0625")      # ... also when it comes back reversed
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    # ---- the reference's ingest, verbatim (trainer.py:98-151) ----
    graph = nx.read_weighted_edgelist(path, nodetype=int)
    n = graph.number_of_nodes()
    adj = nx.adjacency_matrix(graph, nodelist=list(range(n)), weight="weight", dtype=np.float32)
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)
    t = ref_utils.preprocess_adj(adj, is_sparse=True)
    idx = t._indices().numpy()
    np.savez_compressed(os.path.join(HERE, "edges_small_adj.npz"), rows=idx[0], cols=idx[1], vals=t._values().numpy(),
                        n=np.array(n), n_docs=np.array(n_docs), n_topics=np.array(n_topics))
    print("wrote", path, "n =", n, "nnz =", idx.shape[1])


if __name__ == "__main__":
    main()
