"""Workload for the ncu captures of one C3 bench step (evolve L=6 + fused score/rank): two warm steps, then ONE more.
Prints how many library kernels / tcgen05 GEMM launches the warm part issued, so that
    ncu --set full -k regex:gemm_tf32 --launch-skip <gemm launches of the warm part> --launch-count <per step> ...
captures exactly the GEMM launches of the last step.  Run: python profiles/prof_step.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regcn_b200 as R
from regcn_b200 import _lib, evaluate, synth, utils
from bench import build_product_model, model_cfg
case = synth.make_case("c3", 0); n, r = case["num_ents"], case["num_rels"]
model, _ = build_product_model(model_cfg("regcn"), n, r, 0); model = model.cuda()
gl = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
test = torch.from_numpy(case["test"]).cuda(); inv = test[:, [2, 1, 0]].clone(); inv[:, 1] += r
all_t = torch.cat((test, inv)).contiguous()
f = utils.filter_csr_from_snapshot(all_t, 2 * r, 0)
lib = _lib.load()
lib.regcn_prof_enable(1)
for _ in range(2):
    evaluate.evaluate_snapshot(model, gl, all_t, f)
torch.cuda.synchronize()
ms, nl, w = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
lib.regcn_prof_read(0, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(w))
lib.regcn_prof_enable(0)
k0 = lib.regcn_kernel_launches()
evaluate.evaluate_snapshot(model, gl, all_t, f)
torch.cuda.synchronize()
print("warm gemm launches:", nl.value, "kernels in the last step:", lib.regcn_kernel_launches() - k0)
