"""End-to-end loop alone: regcn_b200.test() over a stream of pinned host snapshots (the e2e leg of bench.py), repeated;
prints ms per evaluated timestamp for every repeat.  python profiles/e2e_loop.py [repeats] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import regcn_b200 as R
from regcn_b200 import _lib, synth
from bench import build_product_model, model_cfg
_lib.require_device()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
stream = synth.make_stream("c3", 1000, n_test=8 + steps * (reps + 1))
n, r = stream["num_ents"], stream["num_rels"]
L = len(stream["history"])
model, _ = build_product_model(model_cfg("regcn"), n, r, 0)
model = model.to(dev)
hist = [torch.from_numpy(s).pin_memory() for s in stream["history"]]
tests = [torch.from_numpy(s).pin_memory() for s in stream["tests"]]
R.test(model, hist, tests[:8], r, n, True, test_history_len=L)
pos = 8
for k in range(reps + 1):
    win = (hist + tests[:pos])[-L:]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    R.test(model, win, tests[pos:pos + steps], r, n, True, test_history_len=L)
    b.record()
    torch.cuda.synchronize()
    print(f"repeat {k}: {a.elapsed_time(b) / steps:.3f} ms/step (device events), host wall {1e3 * (time.perf_counter() - t0) / steps:.3f} ms/step")
    pos += steps
