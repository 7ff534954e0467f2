"""Counting GEMM (dot, 3xTF32) vs candidate-table size with an L2 flush between launches: is the per-tile time flat?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regcn_b200 import _lib, ops
lib = _lib.load(); _lib.require_device()
dev = "cuda"; d = 200; B = 8192
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
q = torch.randn(B, d, device=dev, generator=g) * 0.05
qh, ql = ops.split_tf32(q)
for N in (62500, 125000, 250000, 500000, 1000000):
    e = torch.randn(N, d, device=dev, generator=g) * 0.05
    eh, el = ops.split_tf32(e); del e
    tgt = torch.randint(0, N, (B,), device=dev, dtype=torch.int32, generator=g)
    ts = torch.zeros(B, device=dev); raw = torch.zeros(B, device=dev, dtype=torch.int32)
    fn = lambda: _lib.call("regcn_score_count_tf32", qh.data_ptr(), ql.data_ptr(), eh.data_ptr(), el.data_ptr(), B, N, d,
                           ts.data_ptr(), tgt.data_ptr(), raw.data_ptr(), 0, 0, None, None, None, 1.0, None, None, 3)
    for _ in range(2):
        fn()
    out = []
    for fl in (False, True):
        tms = []
        for _ in range(5):
            if fl:
                flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); tms.append(a.elapsed_time(b))
        out.append(sorted(tms)[2])
    tiles = ((B + 127) // 128) * ((N + 255) // 256)
    print(f"N={N:8d} tiles={tiles:7d}  warm {out[0]:7.3f} ms ({out[0] * 1e3 * 148 / tiles:5.2f} us/tile/SM)   flushed {out[1]:7.3f} ms ({out[1] * 1e3 * 148 / tiles:5.2f} us/tile/SM)")
    del eh, el
