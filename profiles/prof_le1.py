"""Workload for the ncu capture of the last-layer GEMM with the fused normalise + time-gate epilogue (Le1 of a C3
snapshot, 23033 x 200 x 200): pre-split A (TMA) and fp32 A (converter warps), 3 launches each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regcn_b200 import _lib, ops
N, d = 23033, 200
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.randn(N, d, device=dev, generator=g) * 0.1
w1 = torch.randn(d, d, device=dev, generator=g) * 0.1
xh, xl = ops.split_tf32(x); w1h, w1l = ops.split_tf32(w1)
skip = torch.full((N,), -1, device=dev, dtype=torch.int32)
o_raw, o_hi, o_lo = (torch.empty(N, d, device=dev) for _ in range(3))
gate = torch.randn(N, d, device=dev, generator=g); gb = torch.randn(d, device=dev, generator=g) * 0.1
hp = torch.randn(N, d, device=dev, generator=g) * 0.1
for _ in range(3):
    _lib.call("regcn_gemm_tf32_layer", xh.data_ptr(), xl.data_ptr(), d, w1h.data_ptr(), w1l.data_ptr(), d, N, d, d, d,
              o_raw.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr(), None, 0, None, skip.data_ptr(), gate.data_ptr(), d,
              gb.data_ptr(), hp.data_ptr(), 1)
for _ in range(3):
    _lib.call("regcn_gemm_tf32_layer_a32", x.data_ptr(), d, d, None, None, 0, 0, None, w1h.data_ptr(), w1l.data_ptr(), d, N,
              d, d, o_raw.data_ptr(), None, None, None, 0, None, skip.data_ptr(), gate.data_ptr(), d, gb.data_ptr(),
              hp.data_ptr(), 1)
torch.cuda.synchronize()
print("ok")
