"""Scoring GEMM with the counting epilogue alone: C3 shape and one C5 entity shard (8192 x 125000), dot and hyperbolic
(RotH form) scores, 3xTF32 / bf16; hyperbolic with the polynomial threshold test on and off.  CUDA events, median of 7."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regcn_b200 import _lib, ops
lib = _lib.load(); _lib.require_device()
dev = "cuda"; d = 200


def med(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


out = []
for shape, B, N in (("c3", 2914, 23033), ("c5 shard 1/8", 8192, 125000), ("c5 full table", 8192, 1000000)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    q = torch.randn(B, d, device=dev, generator=g) * 0.05
    e = torch.randn(N, d, device=dev, generator=g) * 0.05
    tgt = torch.randint(0, N, (B,), device=dev, dtype=torch.int32, generator=g)
    ts = torch.zeros(B, device=dev); raw = torch.zeros(B, device=dev, dtype=torch.int32)
    qh, ql = ops.split_tf32(q); eh, el = ops.split_tf32(e)
    x2, y2 = ops.row_sumsq(q), ops.row_sumsq(e)
    sm = torch.tensor([1.0, 1.0], device=dev)
    for name, hyp, poly in (("dot 3xTF32", 0, 1), ("hyp 3xTF32 exact epilogue", 1, 0), ("hyp 3xTF32 polynomial test", 1, 1)):
        lib.regcn_score_count_poly(poly)
        ms = med(lambda: _lib.call("regcn_score_count_tf32", qh.data_ptr(), ql.data_ptr(), eh.data_ptr(), el.data_ptr(), B, N, d,
                                   ts.data_ptr(), tgt.data_ptr(), raw.data_ptr(), 0, hyp, x2.data_ptr() if hyp else None,
                                   y2.data_ptr() if hyp else None, None, 0.01, sm.data_ptr() if hyp else None, None, 3))
        r = {"shape": shape, "B": B, "N": N, "mode": name, "ms": ms, "algorithmic_tflops": 2.0 * B * N * d / ms / 1e9}
        out.append(r); print(r)
    lib.regcn_score_count_poly(1)
    del q, e, qh, ql, eh, el
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/prof_score.json", "w"), indent=1)
