"""Workload for the ncu captures of one BATCHED bench step (G consecutive ICEWS18-shaped timestamps: shared-trajectory
recurrence, then decode + rank per timestamp -- what bench.py times): two warm batches, then ONE more.  Prints how many
library kernels / tcgen05 GEMM launches the warm part issued, so that
    ncu --set full -k regex:gemm_tf32 --launch-skip <gemm launches of the warm part> --launch-count <per batch> ...
captures exactly the GEMM launches of the last batch.  Run: python profiles/prof_batch.py [G]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regcn_b200 as R
from regcn_b200 import _lib, evaluate, synth, utils
from bench import build_product_model, model_cfg
G = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, r, t, L, tq = synth.SHAPES["c3"]
st = synth.make_stream("c3", 0, n_test=G)
snaps = list(st["history"]) + list(st["tests"][:G - 1])
model, _ = build_product_model(model_cfg("regcn"), n, r, 0); model = model.cuda()
graphs = [R.build_sub_graph(n, r, s_, True, 0) for s_ in snaps]
windows, trips, filts = [], [], []
for g in range(G):
    windows.append(graphs[g:g + L])
    tg = torch.from_numpy(st["tests"][g]).cuda(); ig = tg[:, [2, 1, 0]].clone(); ig[:, 1] += r
    trips.append(torch.cat((tg, ig)).contiguous())
    filts.append(utils.filter_csr_from_snapshot(trips[-1], 2 * r, 0))
lib = _lib.load()
lib.regcn_prof_enable(1)
for _ in range(2):
    evaluate.evaluate_batch(model, windows, trips, filts)
torch.cuda.synchronize()
ms, nl, w = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
lib.regcn_prof_read(0, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(w))
lib.regcn_prof_enable(0)
k0 = lib.regcn_kernel_launches()
evaluate.evaluate_batch(model, windows, trips, filts)
torch.cuda.synchronize()
print("warm gemm launches:", nl.value, "kernels in the last batch:", lib.regcn_kernel_launches() - k0, "timestamps:", G)
