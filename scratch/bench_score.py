"""Fused scoring+count kernel (no score matrix) at the C3 and C5 shapes: time, algorithmic and executed TFLOP/s."""
import sys, json, torch
sys.path.insert(0,'.')
from regcn_b200 import ops, _lib
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
res=[]
for name,B,N in [("c3",2914,23033),("c5 shard 1/8",8192,125000),("c5 full",8192,1000000)]:
    d=200
    q=torch.randn(B,d,device='cuda'); e=torch.randn(N,d,device='cuda')*0.5
    target=torch.randint(0,N,(B,),device='cuda',dtype=torch.int32)
    tscore=torch.zeros(B,device='cuda'); raw=torch.zeros(B,device='cuda',dtype=torch.int32)
    qh,ql=ops.split_tf32(q); eh,el=ops.split_tf32(e)
    qb,eb=ops.to_bf16(q),ops.to_bf16(e)
    for passes in (3,1,0):
        a_,b_=(qb,eb) if passes==0 else (qh,eh)
        def f(): _lib.call("regcn_score_count_tf32", a_.data_ptr(), ql.data_ptr(), b_.data_ptr(), el.data_ptr(), B,N,d, tscore.data_ptr(), target.data_ptr(), raw.data_ptr(), 0, 0, None,None,None, 1.0, None, None, passes)
        ms=timeit(f)
        alg=2.0*B*N*d/ms/1e9
        r=dict(shape=name,B=B,N=N,mode={3:'3xTF32',1:'TF32',0:'bf16'}[passes],ms=ms,algorithmic_TFLOPs=alg,executed_TFLOPs=alg*max(passes,1),frac_of_mode_peak_executed=alg*max(passes,1)/(1369.2 if passes==0 else 1369.2/2),frac_bf16_peak_algorithmic=alg/1369.2)
        print(json.dumps(r)); res.append(r)
    del q,e,qh,ql,eh,el,qb,eb
json.dump(res,open('gpurun_out/score_kernel.json','w'),indent=1)
