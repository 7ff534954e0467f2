"""Warm per-kernel time breakdown of one bench step via torch.profiler (CUPTI): scratch tool, not a bench."""
import sys, json, torch
sys.path.insert(0, '.')
import regcn_b200 as R
from regcn_b200 import ops, synth, utils, evaluate
from bench import build_product_model, model_cfg
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c3"); ap.add_argument("--model", default="regcn")
ap.add_argument("--e2e", action="store_true"); a = ap.parse_args()
case = synth.make_case(a.workload, 0); n, r = case["num_ents"], case["num_rels"]
model, _ = build_product_model(model_cfg(a.model), n, r, 0); model = model.cuda()
gl = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
test = torch.from_numpy(case["test"]).cuda(); inv = test[:, [2, 1, 0]].clone(); inv[:, 1] += r
all_t = torch.cat((test, inv)).contiguous()
f = utils.filter_csr_from_snapshot(all_t, 2 * r, 0)
hh = [torch.from_numpy(s).pin_memory() for s in case["history"]]; th = torch.from_numpy(case["test"]).pin_memory()
dev = torch.device("cuda", 0)
fn = (lambda: evaluate.evaluate_from_host(model, hh, th, n, r, dev)) if a.e2e else (lambda: evaluate.evaluate_snapshot(model, gl, all_t, f))
for _ in range(5): fn()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
K = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(K): fn()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_time_total > 0 and e.device_type.name == "CUDA" or (e.self_device_time_total > 0):
        rows.append((e.self_device_time_total / K, e.count / K, e.key))
rows.sort(reverse=True)
tot = sum(x[0] for x in rows)
print(f"total device us/step {tot:.1f}")
for t, c, k in rows[:40]:
    print(f"{t:9.1f} us  x{c:5.1f}  {t/max(c,1e-9):8.1f} us/launch  {k[:110]}")
