import sys, numpy as np, torch
sys.path.insert(0, '.')
import regcn_b200 as R
from regcn_b200 import optim, synth
from tests.helpers import build_model
cfg = dict(kind="regcn", shape="tiny", seed=0, layer_norm=True)
case = synth.make_case("tiny", 0); n, r = case["num_ents"], case["num_rels"]
m, _ = build_model(cfg, n, r); m = m.cuda()
empty = np.zeros((0, 3), dtype=np.int64)
def run(name, fn):
    try:
        out = fn(); torch.cuda.synchronize(); print("OK  ", name, out)
    except Exception as e:
        print("FAIL", name, type(e).__name__, str(e)[:200])
hist = case["history"]
run("predict, empty middle snapshot", lambda: m.predict([R.build_sub_graph(n, r, s, True, 0) for s in (hist[0], empty, hist[2])], r, None, torch.from_numpy(case["test"]).cuda(), True)[1].shape)
run("predict, single test triple", lambda: m.predict([R.build_sub_graph(n, r, s, True, 0) for s in hist], r, None, torch.from_numpy(case["test"][:1]).cuda(), True)[1].shape)
run("test() with an empty history snapshot", lambda: R.test(m, [hist[0], empty, hist[2]], [case["test"]], r, n, True, test_history_len=3))
run("test() single-triple snapshot", lambda: R.test(m, hist, [case["test"][:1], case["test"]], r, n, True, test_history_len=3))
m.train(); opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
def step(glist, tr):
    l = m.get_loss(glist, torch.from_numpy(tr).cuda(), None, True); (0.7*l[0]+0.3*l[1]+l[2]).backward(); optim.clip_grad_norm_(opt, 1.0); opt.step(); opt.zero_grad(); return float(l[0].detach())
run("train, empty middle snapshot", lambda: step([R.build_sub_graph(n, r, s, True, 0) for s in (hist[0], empty, hist[2])], case["test"]))
run("train, one history snapshot", lambda: step([R.build_sub_graph(n, r, hist[0], True, 0)], case["test"]))
run("train, single triple (B=2)", lambda: step([R.build_sub_graph(n, r, s, True, 0) for s in hist], case["test"][:1]))
run("fit_epoch with an empty snapshot inside", lambda: R.fit_epoch(m, opt, [hist[0], empty, hist[1], hist[2], case["test"]], r, n, 3)["steps"])
