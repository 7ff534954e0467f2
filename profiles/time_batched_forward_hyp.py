"""Device time of one batched recurrence of the hyperbolic model (BASELINE configs[1]: lgcn + RotH, ICEWS14s shape) for G
history windows per recurrence, L2 flushed between repetitions.  python profiles/time_batched_forward_hyp.py [G ...]"""
import json, os, sys
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
n, r, t, L, tq = synth.SHAPES["c1"]
out = {}
for enc in ("hyp_lgcn", "hyp_uv"):
    m, sd = build_product_model(model_cfg(enc), n, r, 0)
    m = m.to(dev).eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for G in Gs:
        rng = np.random.default_rng(0)
        snaps = [synth.make_snapshot(rng, n, r, t, True) for _ in range(L + G - 1)]
        graphs = [R.build_sub_graph(n, r, s, True, 0) for s in snaps]
        windows = [graphs[g:g + L] for g in range(G)]
        ts = []
        for rep in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if G == 1:
                m.forward(windows[0], None, True)
            else:
                m.forward_batch(windows)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        out[f"{enc}_G{G}_us_per_timestamp"] = round(1000 * ts[len(ts) // 2] / G, 1)
print(json.dumps(out))
