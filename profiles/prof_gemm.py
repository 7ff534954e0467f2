"""Workload for the ncu capture of the tcgen05 GEMM: the C3 scoring GEMM with the counting epilogue, the dense scoring
GEMM, and a node GEMM, each launched a few times.  Run: python profiles/prof_gemm.py  (then the same under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regcn_b200 import ops, _lib
ops.set_gemm_impl("tc")
B, N, d = 2914, 23033, 200
q = torch.randn(B, d, device="cuda"); e = torch.randn(N, d, device="cuda") * 0.5
h = torch.randn(N, d, device="cuda"); w = torch.randn(400, d, device="cuda")
target = torch.randint(0, N, (B,), device="cuda", dtype=torch.int32)
fptr = torch.arange(B + 1, device="cuda", dtype=torch.int32); fidx = target.clone()
pa = torch.cat((torch.arange(B, device="cuda"), torch.arange(B, device="cuda"))).to(torch.int32)
pe = torch.cat((target, target)).contiguous()
for _ in range(3):
    ops.gemm(h, w, trans_b=True)                                   # node GEMM  23033 x 400 x 200
    ops.gemm(q, e, trans_b=True)                                   # dense scoring GEMM
    ops.fused_rank_counts(q, e, target, fptr, fidx, pa, pe)        # pair scores + counting epilogue
torch.cuda.synchronize()
print("ok")
