"""Edge kernel at BASELINE configs[4] size (N=1M, E=10M, d=200): uniform and Zipf endpoints, register-staged kernel,
streaming (bulk-copy ring) kernel and the automatic choice.  CUDA events, median of 7."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import regcn_b200 as R
from regcn_b200 import _lib, ops, synth
lib = _lib.load(); _lib.require_device()
dev = "cuda"; n, r, t, d = 1_000_000, 512, 5_000_000, 200
h = torch.randn(n, d, device=dev); rel = torch.randn(2 * r, d, device=dev); o = torch.empty(n, d, device=dev)
alg = 808.0 * 2 * t + 808.0 * n + 800.0 * 2 * r


def med(fn, k=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(k):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[k // 2]


res = []
for name, zipf in (("uniform", False), ("zipf", True)):
    tri = synth.make_snapshot(np.random.default_rng(0), n, r, t, zipf=zipf)
    g = R.build_sub_graph(n, r, tri, True, 0)
    for variant, impl in (("register", 1), ("stream", 3), ("auto", 0)):
        lib.regcn_aggregate_tune(impl)
        ms = med(lambda: ops.union_aggregate(h, rel, g, out=o))
        lib.regcn_aggregate_tune(0)
        rec = {"endpoints": name, "variant": variant, "ms": ms, "GBps": alg / ms / 1e6, "frac_of_6534.8": alg / ms / 1e6 / 6534.8,
               "split_chunks": g.n_split_chunks}
        res.append(rec); print(rec)
    del g
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/prof_edge_zipf.json", "w"), indent=1)
