"""CUPTI timeline of one training step (profiles/prof_train.py's step): device busy time vs span, kernels by total time.
    python profiles/timeline_train.py [shape]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
import bench
import regcn_b200 as R
from regcn_b200 import optim, synth

shape = sys.argv[1] if len(sys.argv) > 1 else "c3"
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

for _ in range(4):
    step()
torch.cuda.synchronize()
hs = []
for _ in range(5):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    hs.append((1e3 * (t1 - t0), 1e3 * (t2 - t0)))
print("host enqueue / to completion ms:", [(round(a, 2), round(b, 2)) for a, b in hs])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "trace_train.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
ev = sorted((e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
os.remove(path)
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy, cs, ce = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > ce:
        busy += ce - cs; cs, ce = s, e
    else:
        ce = max(ce, e)
busy += ce - cs
print(f"{len(ev)} device activities, span {t1 - t0:.0f} us, device busy {busy:.0f} us, idle {t1 - t0 - busy:.0f} us")
agg = {}
for e in ev:
    nm = e["name"].replace("regcn::", "").replace("void ", "").split("(")[0][:70]
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
for nm, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {d:9.1f} us  x{c:<4d} {nm}")
