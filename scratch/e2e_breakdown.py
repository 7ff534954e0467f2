import sys, time, torch, numpy as np
sys.path.insert(0,'.')
import regcn_b200 as R
from regcn_b200 import ops, synth, utils, evaluate
from regcn_b200.graph import build_sub_graphs
from bench import build_product_model, model_cfg
case=synth.make_case("c3",0); n,r=case["num_ents"],case["num_rels"]
model,_=build_product_model(model_cfg("regcn"),n,r,0); model=model.cuda()
hh=[torch.from_numpy(s).pin_memory() for s in case["history"]]; th=torch.from_numpy(case["test"]).pin_memory()
dev=torch.device("cuda",0)
def T(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
print("total e2e ms", T(lambda: evaluate.evaluate_from_host(model,hh,th,n,r,dev)))
print("build graphs ms", T(lambda: build_sub_graphs(n,r,hh,dev)))
gl=build_sub_graphs(n,r,hh,dev)
test=th.to(dev); inv=test[:,[2,1,0]].clone(); inv[:,1]+=r; all_t=torch.cat((test,inv)).contiguous()
print("filter csr ent ms", T(lambda: utils.filter_csr_from_snapshot(all_t,2*r,0,num_answers=n)))
print("filter csr rel ms", T(lambda: utils.filter_csr_from_snapshot(all_t,n,1,num_answers=2*r)))
print("forward ms", T(lambda: model.forward(gl,None,True)))
embs,_,r_emb,_,_=model.forward(gl,None,True); emb=ops.row_map(embs[-1],ops.ROW_NORMALIZE)
f=utils.filter_csr_from_snapshot(all_t,2*r,0,num_answers=n); target=all_t[:,2].to(torch.int32).contiguous()
def fused():
    q,cand,hyp,cb=evaluate._scoring_operands(model,emb,r_emb,all_t); pa,pe=f.pairs(target)
    return ops.fused_rank_counts(q,cand,target,f.ptr,f.idx,pa,pe,hyp=hyp,col_bias=cb)
print("entity query+fused rank ms", T(fused))
print("rel decoder ms", T(lambda: model.rdecoder.forward(emb,r_emb,all_t,mode="test")))
