"""Workload of the hyperbolic training-step capture: optimisation steps of HyperbolicRecurrentRGCN (hyperbolic_uvrgcn +
<decoder>, layer_norm, dropout 0.2) at the C1 (ICEWS14s) shape.  `python profiles/prof_train_hyp.py [decoder] [steps] [encoder]`."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import regcn_b200 as R
from regcn_b200 import optim, synth
from tests.helpers import build_hyp_train_model

decoder = sys.argv[1] if len(sys.argv) > 1 else "hyperbolic_convtranse"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
encoder = sys.argv[3] if len(sys.argv) > 3 else "hyperbolic_uvrgcn"
cfg = dict(kind="hyp", shape="c1", seed=0, encoder=encoder, decoder=decoder, layer_norm=True, gamma=0.15)
case = synth.make_case("c1", 0)
n, r = case["num_ents"], case["num_rels"]
m, _ = build_hyp_train_model(cfg, n, r, dropout=0.2)
m = m.cuda().train()
opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
t = torch.from_numpy(case["test"]).cuda()
for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    le, lr_, ls, lrad = m.get_loss(glist, t, None, True)
    (0.7 * le + 0.3 * lr_ + ls + lrad).backward()
    optim.clip_grad_norm_(opt, 1.0)
    opt.step()
    opt.zero_grad()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{encoder}+{decoder} step {i}: host enqueue {1e3 * (t1 - t0):.2f} ms, to completion {1e3 * (time.perf_counter() - t0):.2f} ms "
          f"loss_e {float(le.detach()):.4f}", flush=True)
