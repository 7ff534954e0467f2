"""Warm per-launch timeline of one bench step (torch.profiler/CUPTI), averaged by position over K steps."""
import sys, json, torch, collections
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
print("n_active", [g.n_active for g in gl], "n_vrows", [g.n_vrows for g in gl], "split", [g.n_split_chunks for g in gl])
test = torch.from_numpy(case["test"]).cuda(); inv = test[:, [2, 1, 0]].clone(); inv[:, 1] += r
all_t = torch.cat((test, inv)).contiguous()
f = utils.filter_csr_from_snapshot(all_t, 2 * r, 0)
hh = [torch.from_numpy(s).pin_memory() for s in case["history"]]; th = torch.from_numpy(case["test"]).pin_memory()
dev = torch.device("cuda", 0)
fn = (lambda: evaluate.evaluate_from_host(model, hh, th, n, r, dev)) if a.e2e else (lambda: evaluate.evaluate_snapshot(model, gl, all_t, f))
for _ in range(5): fn()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
K = 8
seqs = []
for _ in range(K):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
    ev.sort(key=lambda e: e.time_range.start)
    seqs.append([(e.name, e.time_range.end - e.time_range.start, e.time_range.start) for e in ev])
L = min(len(s) for s in seqs)
t0 = [s[0][2] for s in seqs]
tot = 0
for i in range(L):
    nm = seqs[0][i][0]
    d = sum(s[i][1] for s in seqs) / K
    st = sum(s[i][2] - t for s, t in zip(seqs, t0)) / K
    tot += d
    print(f"{i:3d} start {st:8.1f} dur {d:7.1f}  {nm[:90]}")
print("sum of kernel durations", tot, "span", sum(s[L-1][2] + s[L-1][1] - t for s, t in zip(seqs, t0)) / K)
