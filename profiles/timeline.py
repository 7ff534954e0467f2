"""Kernel timeline of the hot path WITHOUT serialisation: torch.profiler (CUPTI activity records) around a few
device-resident steps, so the two streams, programmatic dependent launch and warm caches are all as in the timed
region of bench.py.  Writes gpurun_out/timeline_<tag>.json = [{name, stream, t_us (from the step's first kernel),
dur_us}] for the LAST profiled step and prints a per-kernel summary plus the critical-path view per stream.

    python profiles/timeline.py [c3|c1hyp|c4|score5] [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch
from torch.profiler import ProfilerActivity, profile

import regcn_b200 as R
from regcn_b200 import _lib, evaluate, synth, utils
from bench import build_product_model, model_cfg


def short(name):
    name = name.replace("regcn::", "").replace("void ", "")
    cut = name.find("(")
    return name[:cut] if cut > 0 else name


def run(tag, fn, steps):
    for _ in range(2 if tag == "c5" else 3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    marks = []
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            fn()
            torch.cuda.synchronize()
    path = os.path.join(ROOT, "gpurun_out", f"trace_{tag}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    os.remove(path)
    ev.sort(key=lambda e: e["ts"])
    # split into steps at the L2-flush fill kernels
    groups, cur = [], []
    for e in ev:
        if "fill" in e["name"].lower() and e["dur"] > 20:
            if cur:
                groups.append(cur)
            cur = []
        else:
            cur.append(e)
    if cur:
        groups.append(cur)
    last = groups[-1]
    t0 = last[0]["ts"]
    rows = [{"name": short(e["name"]), "stream": e["args"].get("stream"), "t_us": round(e["ts"] - t0, 2),
             "dur_us": round(e["dur"], 2), "grid": e["args"].get("grid"), "block": e["args"].get("block")}
            for e in last]
    span = max(r["t_us"] + r["dur_us"] for r in rows)
    json.dump({"tag": tag, "span_us": span, "kernels": rows}, open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}.json"), "w"))
    print(f"== {tag}: {len(rows)} kernels, span {span:.1f} us (steps profiled {len(groups)}; spans "
          f"{[round(max(e['ts'] + e['dur'] for e in g) - g[0]['ts'], 1) for g in groups]})")
    agg = {}
    for r in rows:
        a = agg.setdefault((r["name"], r["stream"]), [0, 0.0])
        a[0] += 1
        a[1] += r["dur_us"]
    for (nm, s), (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {d:9.1f} us  x{c:<3d} stream {s}  {nm[:100]}")
    print("  -- first 40 kernels --")
    for r in rows[:(80 if tag == "c5" else 40)]:
        print(f"  t={r['t_us']:8.1f}  dur={r['dur_us']:7.1f}  s={r['stream']}  grid={r['grid']}  {r['name'][:90]}")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    _lib.require_device()
    dev = torch.device("cuda", 0)
    if what in ("c3", "c1", "c4", "c1hyp", "c3hyp"):
        shape = what.replace("hyp", "")
        cfg = model_cfg("hyp_lgcn_roth" if what.endswith("hyp") else "regcn")
        case = synth.make_case(shape, 0)
        n, r = case["num_ents"], case["num_rels"]
        model, _ = build_product_model(cfg, n, r, 0)
        model = model.to(dev)
        gl = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
        test = torch.from_numpy(case["test"]).to(dev)
        inv = test[:, [2, 1, 0]].clone()
        inv[:, 1] += r
        all_t = torch.cat((test, inv)).contiguous()
        f = utils.filter_csr_from_snapshot(all_t, 2 * r, 0)
        run(what, lambda: evaluate.evaluate_snapshot(model, gl, all_t, f), steps)
    elif what.startswith("c3x"):
        # G test timestamps evolved as one block-diagonal graph (profiles/prof_batched_evolve.py)
        import numpy as np
        from prof_batched_evolve import batched_model
        G = int(what[3:])
        n, r, t, L, tq = synth.SHAPES["c3"]
        rng = np.random.default_rng(0)
        snaps = [synth.make_snapshot(rng, n, r, t, True) for _ in range(L + G - 1)]
        m, sd = build_product_model(model_cfg("regcn"), n, r, 0)
        mG = batched_model(m, sd, n, r, G, dev)
        comb = []
        for i in range(L):
            parts = []
            for g in range(G):
                s_ = snaps[g + i].copy()
                s_[:, 0] += g * n
                s_[:, 2] += g * n
                s_[:, 1] += g * r
                parts.append(s_)
            comb.append(R.build_sub_graph(G * n, G * r, np.concatenate(parts), True, 0))
        run(what, lambda: mG.forward(comb, None, True), steps)
    elif what.startswith("c3e"):
        # the whole batched step: G windows evolved together (forward_batch), then scored and ranked one by one
        import numpy as np
        G = int(what[3:])
        n, r, t, L, tq = synth.SHAPES["c3"]
        st = synth.make_stream("c3", 0, n_test=G)
        snaps = list(st["history"]) + list(st["tests"][:G - 1])
        m, sd = build_product_model(model_cfg("regcn"), n, r, 0)
        m = m.to(dev)
        graphs = [R.build_sub_graph(n, r, s_, True, 0) for s_ in snaps]
        windows, trips, filts = [], [], []
        for g in range(G):
            windows.append(graphs[g:g + L])
            tg = torch.from_numpy(st["tests"][g]).to(dev)
            ig = tg[:, [2, 1, 0]].clone()
            ig[:, 1] += r
            trips.append(torch.cat((tg, ig)).contiguous())
            filts.append(utils.filter_csr_from_snapshot(trips[-1], 2 * r, 0))
        run(what, lambda: evaluate.evaluate_batch(m, windows, trips, filts), steps)
    elif what == "c5":
        # BASELINE configs[4]: 1 M entities, 512 relations, 10 M edges per snapshot, L = 3, hyperbolic_uvrgcn + RotH
        import numpy as np
        n, r, t, L5, tq = 1_000_000, 512, 5_000_000, 3, 4096
        rng = np.random.default_rng(5)
        hist = [synth.make_snapshot(rng, n, r, t, zipf=False) for _ in range(L5)]
        test = synth.make_snapshot(rng, n, r, tq, zipf=False)
        m5 = R.HyperbolicRecurrentRGCN("roth", "hyperbolic_uvrgcn", n, r, 0, 0, 200, "sub", 3, num_bases=100,
                                       num_hidden_layers=2, dropout=0.2, c=0.01, self_loop=True, layer_norm=False,
                                       input_dropout=0.2, hidden_dropout=0.2, feat_dropout=0.2, entity_prediction=True,
                                       relation_prediction=True, use_cuda=True, gpu=0, radius_msg_gamma=0.15)
        with torch.no_grad():
            m5.dynamic_emb.copy_(torch.randn(n, 200, generator=torch.Generator().manual_seed(5)) * 0.5)
        m5 = m5.to(dev).eval()
        gl = [R.build_sub_graph(n, r, s_, True, 0) for s_ in hist]
        tt = torch.from_numpy(test).to(dev)
        inv = tt[:, [2, 1, 0]].clone()
        inv[:, 1] += r
        all_t = torch.cat((tt, inv)).contiguous()
        f = utils.filter_csr_from_snapshot(all_t, 2 * r, 0)
        run(what, lambda: evaluate.evaluate_snapshot(m5, gl, all_t, f), min(steps, 2))
    else:
        raise SystemExit("unknown workload " + what)


if __name__ == "__main__":
    main()
