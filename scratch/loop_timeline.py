"""GPU timeline (CUPTI) of the pipelined test() loop: prints per-step GPU busy time and idle gaps."""
import sys, torch
sys.path.insert(0,'.')
import regcn_b200 as R
from regcn_b200 import synth
from bench import build_product_model, model_cfg
st=synth.make_stream("c3",1000,n_test=16); n,r=st["num_ents"],st["num_rels"]
model,_=build_product_model(model_cfg("regcn"),n,r,0); model=model.cuda()
h=[torch.from_numpy(s).pin_memory() for s in st["history"]]; t=[torch.from_numpy(s).pin_memory() for s in st["tests"]]
R.test(model,h,t[:8],r,n,True,test_history_len=6)
win=(h+t[:8])[-6:]
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    R.test(model,win,t[8:],r,n,True,test_history_len=6); torch.cuda.synchronize()
ev=[e for e in prof.events() if e.device_type.name=="CUDA"]
ev.sort(key=lambda e:e.time_range.start)
t0=ev[0].time_range.start; end=max(e.time_range.end for e in ev)
# merge intervals for busy time
busy=0; cur_s=None; cur_e=None
gaps=[]
for e in ev:
    s_,e_=e.time_range.start,e.time_range.end
    if cur_e is None: cur_s,cur_e=s_,e_
    elif s_<=cur_e: cur_e=max(cur_e,e_)
    else:
        busy+=cur_e-cur_s; gaps.append((s_-cur_e, cur_e-t0, e.name[:60])); cur_s,cur_e=s_,e_
busy+=cur_e-cur_s
print("span us", end-t0, "busy us", busy, "steps 8 -> per step span", (end-t0)/8, "busy", busy/8)
gaps.sort(reverse=True)
for g in gaps[:25]: print(f"gap {g[0]:8.1f} us at {g[1]:9.1f} before {g[2]}")
