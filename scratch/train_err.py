import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import regcn_b200 as R
from oracle import restate, synth
from regcn_b200 import optim
from tests.helpers import sample_of
z = np.load("tests/golden/train_regcn.npz")
for name, shape, seed, ln in [("regcn_tiny_s1_noln", "tiny", 1, False), ("regcn_small_s2", "small", 2, True)]:
    case = synth.make_case(shape, seed)
    n, r = case["num_ents"], case["num_rels"]
    m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, 200, "sub", 3, num_bases=100, num_basis=-1,
                        num_hidden_layers=2, dropout=0.0, self_loop=True, skip_connect=False, layer_norm=ln,
                        input_dropout=0.0, hidden_dropout=0.0, feat_dropout=0.0, entity_prediction=True,
                        relation_prediction=True, use_cuda=True, gpu=0)
    sd = synth.fill_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    m = m.cuda().train()
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    le, lr_, ls = m.get_loss(glist, torch.from_numpy(case["test"]).cuda(), None, True)
    (0.7 * le + 0.3 * lr_).backward()
    graphs = [restate.build_edges(s, n, r) for s in case["history"]]
    log, _ = restate.regcn_train_steps(sd, graphs, r, case["test"], layer_norm=ln, steps=1, dtype=torch.float64)
    tn = float(z[f"{name}.s0.grad_norm"])
    print(name, "grad norm", tn)
    for k, p in m.named_parameters():
        if p.grad is None: continue
        t = sample_of(log[0]["grads"][k].numpy().astype(np.float64)).astype(np.float64)
        t = log[0]["grads"][k].numpy().reshape(-1)
        step = max(1, t.size // 1024); t = t[::step][:1024]
        g = p.grad.detach().cpu().numpy().reshape(-1)[::step][:1024].astype(np.float64)
        ref = z[f"{name}.s0.g.{k}"].astype(np.float64)
        mx = np.abs(t).max()
        print(f"  {k:40s} max|g| {mx:9.3e}  kernel err {np.abs(g-t).max()/max(mx,1e-3*tn):9.2e}  ref err {np.abs(ref-t).max()/max(mx,1e-3*tn):9.2e}")
