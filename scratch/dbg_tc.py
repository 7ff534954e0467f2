import torch, sys
sys.path.insert(0, '.')
from regcn_b200 import ops
ops.set_gemm_impl("tc")
def run(M,N,K,trans_b,bias,impl="tc"):
    ops.set_gemm_impl(impl)
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M,K,generator=g); b = torch.randn((N,K) if trans_b else (K,N), generator=g)
    bv = torch.randn(N,generator=g) if bias else None
    ref = a.double() @ (b.double().t() if trans_b else b.double())
    if bias: ref = ref + bv.double()
    out = ops.gemm(a.cuda(), b.cuda(), trans_b=trans_b, bias=bv.cuda() if bias else None)
    err = (out.cpu().double()-ref).abs()
    bad = err > 5e-4*torch.clamp(ref.abs(),min=1.0)
    rows = bad.any(1).nonzero().view(-1); cols = bad.any(0).nonzero().view(-1)
    print(f"M{M} N{N} K{K} tb{trans_b} bias{bias} {impl}: maxerr {err.max():.3e} meanerr {err.mean():.3e} nbad {int(bad.sum())}",
          "rows", rows[:8].tolist(), "..", rows[-3:].tolist(), "cols", cols[:8].tolist(), "..", cols[-3:].tolist())
for cfg in [(300,600,400,True,True),(300,600,400,True,False),(300,600,400,False,False),(300,600,200,True,False),(300,208,400,True,False),
            (300,256,400,True,False),(128,208,400,True,False),(128,208,224,True,False),(128,208,256,True,False),(128,64,416,True,False),(128,64,1024,True,False),
            (128,64,96,True,False),(128,64,128,True,False)]:
    run(*cfg)
run(300,600,400,True,False,"tc1")
run(128,64,1024,True,False,"tc1")
