"""CPU oracle of the TopicGCN graph-convolution hot path — TEST INFRASTRUCTURE, never imported by the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may use this package.
See oracle/gcn_oracle.py and oracle/spmm_oracle.c for what is restated and how it is pinned.
"""
