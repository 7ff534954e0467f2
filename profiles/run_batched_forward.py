"""Two batched recurrences (G windows of the ICEWS18 shape as one block-diagonal graph) and nothing else: the command
ncu wraps to capture the evolution GEMMs at their batched size.   python profiles/run_batched_forward.py [G] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import numpy as np
import torch

import regcn_b200 as R
from regcn_b200 import _lib, synth
from bench import build_product_model, model_cfg

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_lib.require_device()
dev = torch.device("cuda", 0)
n, r, t, L, tq = synth.SHAPES["c3"]
rng = np.random.default_rng(0)
snaps = [synth.make_snapshot(rng, n, r, t, True) for _ in range(L + G - 1)]
m, sd = build_product_model(model_cfg("regcn"), n, r, 0)
m = m.to(dev)
graphs = [R.build_sub_graph(n, r, s, True, 0) for s in snaps]
windows = [graphs[g:g + L] for g in range(G)]
for _ in range(reps):
    out = m.forward_batch(windows)
torch.cuda.synchronize()
print("ok", float(out[0][0].abs().sum()))
