"""Workload for the ncu capture of the edge kernel (K4 union aggregate) at BASELINE configs[4] size (N = 1M, E = 10M,
d = 200): register variant (impl 1), per-row bulk variant (2), streaming variant (3); uniform endpoints by default,
`zipf` as argv[1] for the hub-skewed graph.  Run: python profiles/prof_edge_stream.py [zipf]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import regcn_b200 as R
from regcn_b200 import ops, synth
n, r, t, d = 1_000_000, 512, 5_000_000, 200
zipf = len(sys.argv) > 1 and sys.argv[1] == "zipf"
tri = synth.make_snapshot(np.random.default_rng(0), n, r, t, zipf=zipf)
g = R.build_sub_graph(n, r, tri, True, 0)
h = torch.randn(n, d, device="cuda"); rel = torch.randn(2 * r, d, device="cuda"); o = torch.empty(n, d, device="cuda")
for impl in (1, 2, 3):
    R._lib.load().regcn_aggregate_tune(impl)
    for _ in range(2):
        ops.union_aggregate(h, rel, g, out=o)
torch.cuda.synchronize()
print("ok", g.n_vrows, g.n_split_chunks)
