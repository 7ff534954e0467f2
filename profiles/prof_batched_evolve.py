"""Does evolving G test timestamps at once pay?  The windows of G consecutive test timestamps are independent
(src/main.py:60-90: every timestamp re-runs the recurrence over its own L history snapshots), so they can be evolved as
ONE block-diagonal graph of G*N nodes and G*R relations per recurrence step: the same kernels, G times the rows per
launch.  This script checks that the rows of the batched run equal the per-timestamp runs and times both.

    python profiles/prof_batched_evolve.py [c3] [reps]
"""
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


def batched_model(m, sd, n, r, G, dev):
    cfg = model_cfg("regcn")
    mG, _ = build_product_model(cfg, G * n, G * r, 0)
    sdG = mG.state_dict()
    for k, v in sd.items():
        if k == "dynamic_emb":
            sdG[k] = v.repeat(G, 1)
        elif k in ("emb_rel", "rgcn.rel_emb"):
            sdG[k] = torch.cat([v[:r]] * G + [v[r:]] * G)
        elif sdG[k].shape == v.shape:
            sdG[k] = v
    mG.load_state_dict(sdG)
    return mG.to(dev).eval()


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "c3"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    _lib.require_device()
    dev = torch.device("cuda", 0)
    n, r, t, L, tq = synth.SHAPES[shape]
    rng = np.random.default_rng(0)
    GMAX = 16
    snaps = [synth.make_snapshot(rng, n, r, t, True) for _ in range(L + GMAX - 1)]
    m, sd = build_product_model(model_cfg("regcn"), n, r, 0)
    m = m.to(dev)
    graphs = [R.build_sub_graph(n, r, s, True, 0) for s in snaps]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    out = {"shape": shape, "L": L}
    single = [m.forward(graphs[g:g + L], None, True) for g in range(GMAX)]
    single_h = [s[0][-1].clone() for s in single]
    single_r = [s[2].clone() for s in single]
    t1 = timed(lambda: m.forward(graphs[0:L], None, True))
    out["per_timestamp_ms"] = t1
    print(f"G=1: {t1:.3f} ms per timestamp", flush=True)
    for G in (2, 4, 8, 16):
        mG = batched_model(m, sd, n, r, G, dev)
        comb = []
        for i in range(L):
            parts = []
            for g in range(G):
                s = snaps[g + i].copy()
                s[:, 0] += g * n
                s[:, 2] += g * n
                s[:, 1] += g * r
                parts.append(s)
            comb.append(R.build_sub_graph(G * n, G * r, np.concatenate(parts), True, 0))
        embs, _, h0, _, _ = mG.forward(comb, None, True)
        hG = embs[-1]
        dh = max(float((hG[g * n:(g + 1) * n] - single_h[g]).abs().max()) for g in range(G))
        dr = max(float((torch.cat((h0[g * r:(g + 1) * r], h0[G * r + g * r:G * r + (g + 1) * r])) - single_r[g]).abs().max())
                 for g in range(G))
        tg = timed(lambda: mG.forward(comb, None, True))
        out[f"G{G}"] = {"ms_per_batch": tg, "ms_per_timestamp": tg / G, "max_abs_diff_h": dh, "max_abs_diff_rel": dr}
        print(f"G={G}: {tg:.3f} ms per batch = {tg / G:.3f} per timestamp; max |dh| {dh:.3e} max |dr| {dr:.3e}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "prof_batched_evolve.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
