"""ncu workload: one warm forward (so weight splits / static GEMMs are cached) then ONE more forward of the c3 model."""
import sys, ctypes, torch
sys.path.insert(0, '.')
import regcn_b200 as R
from regcn_b200 import synth, _lib
from bench import build_product_model, model_cfg
case = synth.make_case("c3", 0); n, r = case["num_ents"], case["num_rels"]
model, _ = build_product_model(model_cfg("regcn"), n, r, 0); model = model.cuda()
gl = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
lib = _lib.load()
lib.regcn_prof_enable(1)
model.forward(gl, None, True); torch.cuda.synchronize()
ms, nl, w = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
lib.regcn_prof_read(0, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(w))
lib.regcn_prof_enable(0)
print("gemm launches in the warm forward:", nl.value)
model.forward(gl, None, True); torch.cuda.synchronize()
