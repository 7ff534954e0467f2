"""Workload for the ncu capture of the edge kernel (K4 union aggregate) at an HBM-bound size: 1M entities,
10M edges, d=200 (BASELINE.json configs[4] shape), Zipf endpoints.  Run: python profiles/prof_edge.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import regcn_b200 as R
from regcn_b200 import ops, synth
n, r, t, d = 1_000_000, 512, 5_000_000, 200
tri = synth.make_snapshot(np.random.default_rng(0), n, r, t, zipf=True)
g = R.build_sub_graph(n, r, tri, True, 0)
h = torch.randn(n, d, device="cuda"); rel = torch.randn(2 * r, d, device="cuda"); o = torch.empty(n, d, device="cuda")
for _ in range(3):
    ops.union_aggregate(h, rel, g, out=o)
torch.cuda.synchronize()
print("ok", g.n_vrows, g.n_split_chunks)
