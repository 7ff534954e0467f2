import sys, time, torch
sys.path.insert(0,'.')
import regcn_b200 as R
from regcn_b200 import synth
from bench import build_product_model, model_cfg
st=synth.make_stream("c3",1000,n_test=14); n,r=st["num_ents"],st["num_rels"]
model,_=build_product_model(model_cfg("regcn"),n,r,0); model=model.cuda()
h=[torch.from_numpy(s).pin_memory() for s in st["history"]]; t=[torch.from_numpy(s).pin_memory() for s in st["tests"]]
R.test(model,h,t[:8],r,n,True,test_history_len=6)
win=(h+t[:8])[-6:]
torch.cuda.synchronize(); t0=time.perf_counter()
R.test(model,win,t[8:],r,n,True,test_history_len=6)
torch.cuda.synchronize(); print("ms/step", (time.perf_counter()-t0)/6*1e3)
