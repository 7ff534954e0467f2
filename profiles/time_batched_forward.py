"""Device time of one batched recurrence (G history windows of the ICEWS18 shape) with the shared-trajectory engine and
with the full block-diagonal recurrence, L2 flushed between repetitions.
   python profiles/time_batched_forward.py [G ...]       (REGCN_SHARED_SIDE_SMS etc. from the environment)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import regcn_b200 as R
from regcn_b200 import _lib, synth
from bench import build_product_model, model_cfg

Gs = [int(a) for a in sys.argv[1:]] or [8]
_lib.require_device()
dev = torch.device("cuda", 0)
n, r, t, L, tq = synth.SHAPES["c3"]
m, sd = build_product_model(model_cfg("regcn"), n, r, 0)
m = m.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for G in Gs:
    rng = np.random.default_rng(0)
    snaps = [synth.make_snapshot(rng, n, r, t, True) for _ in range(L + G - 1)]
    graphs = [R.build_sub_graph(n, r, s, True, 0) for s in snaps]
    windows = [graphs[g:g + L] for g in range(G)]
    for mode in ("1", "0"):
        os.environ["REGCN_SHARED_ROWS"] = mode
        ts = []
        for rep in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = m.forward_batch(windows)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        out[f"G{G}_{'shared' if mode == '1' else 'full'}_us_per_timestamp"] = round(1000 * ts[len(ts) // 2] / G, 1)
        if os.environ.get("ONLY_SHARED") == "1":
            break
print(json.dumps(out))
