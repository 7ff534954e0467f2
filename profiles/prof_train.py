"""Workload of the training-step captures: one optimisation step of the C3 (ICEWS18-shaped) RE-GCN
(get_loss in train() mode -> backward -> clip -> Adam).  `python profiles/prof_train.py [shape] [steps]`."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import regcn_b200 as R
from regcn_b200 import optim, synth

shape = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = bench.model_cfg("regcn")
case = synth.make_case(shape, 0)
n, r = case["num_ents"], case["num_rels"]
m, _ = bench.build_product_model(cfg, n, r, 0)
m = m.cuda().train()
opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
t = torch.from_numpy(case["test"]).cuda()


def step():
    le, lr_, ls = m.get_loss(glist, t, None, True)
    (0.7 * le + 0.3 * lr_ + ls).backward()
    optim.clip_grad_norm_(opt, 1.0)
    opt.step()
    opt.zero_grad()


for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: host enqueue {1e3 * (t1 - t0):.2f} ms, to completion {1e3 * (t2 - t0):.2f} ms", flush=True)
