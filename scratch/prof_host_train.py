import os, sys, cProfile, pstats
sys.path.insert(0, '.')
import torch
import regcn_b200 as R
from regcn_b200 import optim, synth
from tests.helpers import build_hyp_train_model
cfg = dict(kind="hyp", shape="c1", seed=0, encoder="hyperbolic_uvrgcn", decoder="hyperbolic_convtranse", layer_norm=True, gamma=0.15)
case = synth.make_case("c1", 0); n, r = case["num_ents"], case["num_rels"]
m, _ = build_hyp_train_model(cfg, n, r, dropout=0.2); m = m.cuda().train()
opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
t = torch.from_numpy(case["test"]).cuda()
def step():
    le, lr_, ls, lrad = m.get_loss(glist, t, None, True)
    (0.7 * le + 0.3 * lr_ + ls + lrad).backward()
    optim.clip_grad_norm_(opt, 1.0); opt.step(); opt.zero_grad()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
