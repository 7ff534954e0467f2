"""CUPTI kernel timeline of the end-to-end loop regcn_b200.test() (pinned host snapshots): per-step span, busy time per
stream and idle gaps.  python profiles/timeline_e2e.py [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import regcn_b200 as R
from regcn_b200 import _lib, synth
from bench import build_product_model, model_cfg
_lib.require_device()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
stream = synth.make_stream("c3", 1000, n_test=8 + 3 * steps)
n, r = stream["num_ents"], stream["num_rels"]; L = len(stream["history"])
model, _ = build_product_model(model_cfg("regcn"), n, r, 0); model = model.to(dev)
hist = [torch.from_numpy(s).pin_memory() for s in stream["history"]]
tests = [torch.from_numpy(s).pin_memory() for s in stream["tests"]]
R.test(model, hist, tests[:8], r, n, True, test_history_len=L)
pos = 8
for _ in range(2):
    R.test(model, (hist + tests[:pos])[-L:], tests[pos:pos + steps], r, n, True, test_history_len=L); pos += steps
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    R.test(model, (hist + tests[:pos])[-L:], tests[pos:pos + steps], r, n, True, test_history_len=L)
    torch.cuda.synchronize()
path = "gpurun_out/trace_e2e.json"; os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace(path)
tr = json.load(open(path))["traceEvents"]; os.remove(path)
ev = sorted((e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} device activities over {t1 - t0:.0f} us = {(t1 - t0) / steps:.0f} us per step ({len(ev) / steps:.0f} per step)")
# union busy time
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; gaps.append((s - cur_e, cur_e - t0)); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"device busy (any stream) {busy / steps:.0f} us per step, idle {(t1 - t0 - busy) / steps:.0f} us per step")
gaps.sort(reverse=True)
print("largest idle gaps (us, at):", [(round(g, 1), round(a)) for g, a in gaps[:12]])
json.dump([{"t": round(e["ts"] - t0, 1), "dur": round(e["dur"], 1), "s": e["args"].get("stream"), "name": e["name"].replace("regcn::", "").split("(")[0][:60]}
           for e in ev], open("gpurun_out/timeline_e2e.json", "w"))
agg = {}
for e in ev:
    nm = e["name"].replace("regcn::", "").split("(")[0][:60]
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
for nm, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {d / steps:8.1f} us/step  x{c / steps:5.1f}  {nm}")
# one step in detail: activities between the 3rd and 4th csr_build
marks = [e["ts"] for e in ev if "csr_build" in e["name"]]
if len(marks) > 4:
    a, b = marks[3], marks[4]
    for e in ev:
        if a <= e["ts"] < b:
            print(f"t={e['ts'] - a:8.1f} dur={e['dur']:7.1f} s={e['args'].get('stream')} {e['name'].replace('regcn::', '')[:70]}")
