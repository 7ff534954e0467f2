"""Edge-kernel (K4 union aggregate) stress: achieved algorithmic GB/s at HBM-bound sizes (SURVEY 8d formula)."""
import sys, json, numpy as np, torch
sys.path.insert(0, '.')
import regcn_b200 as R
from regcn_b200 import ops, synth
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
out=[]
for name,n,r,t,zipf in [("c4d GDELT-dense",7691,240,50000,True),("200k/2M zipf",200_000,256,1_000_000,True),("c5 1M/10M zipf",1_000_000,512,5_000_000,True),("c5 1M/10M uniform",1_000_000,512,5_000_000,False)]:
    rng=np.random.default_rng(0)
    tri=synth.make_snapshot(rng,n,r,t,zipf=zipf)
    g=R.build_sub_graph(n,r,tri,True,0)
    d=200
    h=torch.randn(n,d,device='cuda'); rel=torch.randn(2*r,d,device='cuda')
    o=torch.empty(n,d,device='cuda')
    E=2*t
    bytes_=808.0*E+808.0*n+800.0*2*r
    for impl in (1,2,3):
        R._lib.load().regcn_aggregate_tune(impl)
        ms=timeit(lambda: ops.union_aggregate(h,rel,g,out=o))
        res=dict(case=name,impl={1:"registers",2:"cp.async.bulk",3:"stream"}[impl],N=n,E=E,max_deg=g.max_hub_degree,n_vrows=g.n_vrows,n_split=g.n_split_chunks,ms=ms,GBs=bytes_/ms/1e6,frac=bytes_/ms/1e6/peak)
        print(json.dumps(res)); out.append(res)
    R._lib.load().regcn_aggregate_tune(0)
    del g,h,o
json.dump(out, open('gpurun_out/edge_stress.json','w'), indent=1)
