import sys, torch
sys.path.insert(0,'.')
from regcn_b200 import ops, _lib
lib=_lib.load()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n*1e3
shapes=[("node P",23033,200,200),("node L1",23033,400,200),("node L0",23033,600,200),("gru",512,600,200),("score",2914,23033,200),("fc",2914,200,10000)]
for name,M,N,K in shapes:
    a=torch.randn(M,K,device='cuda'); b=torch.randn(N,K,device='cuda')
    ahi,alo=ops.split_tf32(a); bhi,blo=ops.split_tf32(b)
    out=torch.empty(M,(N+3)//4*4,device='cuda')
    for passes in (3,1):
      for bn,st in [(0,0),(256,0),(208,0),(128,0),(112,0),(64,0),(208,1),(128,2),(64,2)]:
        if bn and bn> ((N+15)//16*16): continue
        lib.regcn_gemm_tf32_tune(bn,st)
        sk = 6 if name=="fc" else 1
        ws=torch.empty(sk*M*N,device='cuda') if sk>1 else None
        def f(): _lib.call("regcn_gemm_tf32", ahi.data_ptr(), alo.data_ptr(), K, bhi.data_ptr(), blo.data_ptr(), K, out.data_ptr(), out.stride(0), M,N,K, None, 0, passes, sk, ws.data_ptr() if ws is not None else None, ws.numel()*4 if ws is not None else 0)
        t=timeit(f)
        print(f"{name:8s} M{M} N{N} K{K} passes{passes} bn{bn:3d} st{st}: {t:8.1f} us  {2*M*N*K/t/1e6:7.1f} TF/s algorithmic")
lib.regcn_gemm_tf32_tune(0,0)
