#!/usr/bin/env python
"""Benchmark of the RE-GCN hot path on B200: one "step" = one evaluated test timestamp of the ICEWS18-shaped
workload (BASELINE.json configs[2]): evolve L=6 history snapshots, score all B=2914 queries against all
N=23033 entities, rank raw + time-filtered.

    python bench.py --gpus N --steps K --warmup W                 # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W  # CPU arm: the oracle port on the host cores

Prints ONE JSON line (rank 0).  `value` is all-entity-ranked queries/s with inputs resident in HBM; `e2e` is the
same metric through the public API from pinned HOST buffers (H2D of the triples, device edge-index build, predict,
ranking, D2H of the ranks); `snapshot_steps_per_s` is the other half of BASELINE.json's metric (evolution only).
N > 1: every rank evaluates its own test timestamps (independent units, weak scaling, no data-path collective);
the entity-sharded scoring + rank-merge path (NCCL all_reduce of counts) is timed separately and reported under
"entity_sharded".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "c3"
H_DIM, N_BASES, N_LAYERS = 200, 100, 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--model", default="regcn", choices=["regcn", "hyp_lgcn_roth", "hyp_uv_roth"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the HBM-bound edge-kernel and the scoring-kernel sections")
    ap.add_argument("--gemm", default=None, help="tc (3xTF32, default) | tc1 (single TF32 pass, reported separately)")
    return ap.parse_args()


def model_cfg(name):
    if name == "regcn":
        return dict(kind="regcn", layer_norm=True)
    enc = "lgcn" if "lgcn" in name else "hyperbolic_uvrgcn"
    return dict(kind="hyp", layer_norm=False, encoder=enc, decoder="roth", gamma=0.15)


def build_product_model(cfg, n, r, seed):
    import regcn_b200 as R
    from regcn_b200 import synth
    if cfg["kind"] == "regcn":
        m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, H_DIM, "sub", 3, num_bases=N_BASES, num_basis=-1,
                            num_hidden_layers=N_LAYERS, dropout=0.2, self_loop=True, skip_connect=False,
                            layer_norm=cfg["layer_norm"], input_dropout=0.2, hidden_dropout=0.2, feat_dropout=0.2,
                            entity_prediction=True, relation_prediction=True, use_cuda=True, gpu=0)
    else:
        m = R.HyperbolicRecurrentRGCN(cfg["decoder"], cfg["encoder"], n, r, 0, 0, H_DIM, "sub", 3, num_bases=N_BASES,
                                      num_hidden_layers=N_LAYERS, dropout=0.2, c=0.01, self_loop=True,
                                      layer_norm=cfg["layer_norm"], input_dropout=0.2, hidden_dropout=0.2,
                                      feat_dropout=0.2, entity_prediction=True, relation_prediction=True,
                                      use_cuda=True, gpu=0, radius_msg_gamma=cfg["gamma"])
    sd = synth.fill_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    return m.eval(), sd


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_step(sd, graphs, r, test, cfg):
    """One step of the same workload on the host cores through the oracle port (oracle/restate.py)."""
    from oracle import restate
    from regcn_b200 import synth
    import torch
    with torch.no_grad():
        if cfg["kind"] == "regcn":
            all_t, score, _, _, _ = restate.regcn_predict(sd, graphs, r, test, layer_norm=cfg["layer_norm"])
        else:
            all_t, score, _, _, _ = restate.hyp_predict(sd, graphs, r, test, c=0.01, decoder=cfg["decoder"],
                                                        encoder=cfg["encoder"], gamma=cfg["gamma"],
                                                        num_bases=min(N_BASES, 2 * r))
    all_ans = synth.answers_of(test, r, False)
    return restate.total_rank(all_t, score.numpy(), all_ans, 0)


def run_cpu_arm(args, steps, warmup, quiet=False):
    """The reference's evaluation loop (src/main.py:33-123) through the oracle port on the host cores: per step rebuild
    the L history graphs, predict, rank entities raw + filtered; the window slides like the GPU arm's."""
    import torch
    from oracle import restate
    from regcn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = model_cfg(args.model)
    stream = synth.make_stream(args.workload, 1000, n_test=warmup + steps)
    n, r = stream["num_ents"], stream["num_rels"]
    _, sd = build_product_model(cfg, n, r, 0)
    window = list(stream["history"])
    B = 2 * len(stream["tests"][0])
    t0 = None
    for k, snap in enumerate(stream["tests"]):
        if k == warmup:
            t0 = time.perf_counter()
        graphs = [restate.build_edges(s, n, r) for s in window]
        cpu_port_step(sd, graphs, r, snap, cfg)
        window.pop(0)
        window.append(snap)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": B / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of the sliding-window evaluation loop on workload {args.workload} (per step: "
                      f"rebuild L={len(window)} graphs, evolve, score {B}x{n}, raw/filtered rank) after {warmup} warm-up, "
                      f"oracle/restate.py with torch CPU ops on {cores} threads; scatter-sum is index_add_, not DGL's kernel",
            "ms_per_step": dt * 1e3}


def run_stress(dev, hbm_peak, tf_peak):
    """(1) K4 union aggregate at BASELINE configs[4] size (N = 1M entities, E = 10M edges, d = 200): algorithmic bytes
    808*E + 808*N + 800*2R over the CUDA-event time, uniform and Zipf endpoints.  (2) The fused
    scoring + count GEMM alone at the C3 shape and at one C5 entity shard, fp32-parity (3xTF32) and bf16 modes."""
    import numpy as np
    import torch
    import regcn_b200 as R
    from regcn_b200 import _lib, ops, synth

    def med(fn, n=7):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    n, r, t, d = 1_000_000, 512, 5_000_000, 200
    h = torch.randn(n, d, device=dev)
    rel = torch.randn(2 * r, d, device=dev)
    o = torch.empty(n, d, device=dev)
    edge = {"kernel": "regcn::union_aggregate_{stream,}kernel (+ radix-32 fix-up for hub rows)", "bound": "hbm",
            "peak": hbm_peak, "unit": "GB/s", "shape": f"N={n} E={2 * t} d={d} 2R={2 * r} (BASELINE configs[4])",
            "algorithmic_bytes_per_launch": 808.0 * 2 * t + 808.0 * n + 800.0 * 2 * r}
    for name, zipf in (("uniform", False), ("zipf", True)):
        tri = synth.make_snapshot(np.random.default_rng(0), n, r, t, zipf=zipf)
        g = R.build_sub_graph(n, r, tri, True, dev.index or 0)
        ms = med(lambda: ops.union_aggregate(h, rel, g, out=o))
        gbs = edge["algorithmic_bytes_per_launch"] / ms / 1e6
        edge[name] = {"ms": ms, "achieved": gbs, "frac": gbs / hbm_peak, "split_chunks": g.n_split_chunks}
        del g
    edge["achieved"], edge["frac"] = edge["uniform"]["achieved"], edge["uniform"]["frac"]
    out["edge_kernel_hbm_bound"] = edge
    del h, o
    score = {"kernel": "regcn::tc::gemm_tf32_kernel<1> (counting epilogue, no score matrix)", "bound": "tensor",
             "peak": tf_peak, "unit": "TFLOP/s", "cases": []}
    for shape, B, N in (("c3", 2914, 23033), ("c5 shard 1/8", 8192, 125000)):
        q = torch.randn(B, d, device=dev)
        e = torch.randn(N, d, device=dev) * 0.5
        target = torch.randint(0, N, (B,), device=dev, dtype=torch.int32)
        tscore = torch.zeros(B, device=dev)
        raw = torch.zeros(B, device=dev, dtype=torch.int32)
        qh, ql = ops.split_tf32(q)
        eh, el = ops.split_tf32(e)
        qb, eb = ops.to_bf16(q), ops.to_bf16(e)
        for mode, passes in (("3xTF32 (fp32 parity)", 3), ("bf16", 0)):
            a_, b_ = (qb, eb) if passes == 0 else (qh, eh)
            ms = med(lambda: _lib.call("regcn_score_count_tf32", a_.data_ptr(), ql.data_ptr(), b_.data_ptr(),
                                       el.data_ptr(), B, N, d, tscore.data_ptr(), target.data_ptr(), raw.data_ptr(), 0, 0,
                                       None, None, None, 1.0, None, None, passes))
            alg = 2.0 * B * N * d / ms / 1e9
            ex = alg * max(passes, 1)
            mode_peak = tf_peak if passes == 0 else tf_peak / 2      # TF32 dense peak = half the bf16 figure
            score["cases"].append({"shape": shape, "B": B, "N": N, "mode": mode, "ms": ms, "algorithmic": alg,
                                   "executed": ex, "frac_of_mode_peak_executed": ex / mode_peak,
                                   "frac_of_bf16_peak_algorithmic": alg / tf_peak})
        del q, e, qh, ql, eh, el, qb, eb
    out["scoring_kernel"] = score
    return out


def run_cpu_train_arm(args, steps=1, warmup=1):
    """The reference's optimisation step (src/main.py:233-246: get_loss in train() mode, backward, clip_grad_norm_,
    Adam) through the oracle port (torch autograd over oracle/restate.py) on the host cores; dropout 0 (mask draws are
    negligible next to the GEMMs)."""
    import torch
    from oracle import restate
    from regcn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = model_cfg(args.model)
    case = synth.make_case(args.workload, 0)
    n, r = case["num_ents"], case["num_rels"]
    _, sd = build_product_model(cfg, n, r, 0)
    graphs = [restate.build_edges(s, n, r) for s in case["history"]]
    restate.regcn_train_steps(sd, graphs, r, case["test"], layer_norm=cfg["layer_norm"], steps=warmup)
    t0 = time.perf_counter()
    restate.regcn_train_steps(sd, graphs, r, case["test"], layer_norm=cfg["layer_norm"], steps=steps)
    dt = (time.perf_counter() - t0) / steps
    L = len(graphs)
    return {"value": L / dt, "unit": "train snapshot-steps/s", "ms_per_step": dt * 1e3, "cores": cores, "kind": "port",
            "sample": f"{steps} optimisation step(s) on workload {args.workload} (evolve L={L} snapshots, ConvTransE + "
                      f"ConvTransR heads over all candidates, backward, clip, Adam) after {warmup} warm-up; torch autograd "
                      f"over oracle/restate.py on {cores} threads"}


def workload_string(args):
    """The same `config.workload` text for both arms."""
    from regcn_b200 import synth
    n, r, t, hist, tq = synth.SHAPES[args.workload]
    if args.model == "regcn":
        return (f"{args.workload}: ICEWS18-shaped N={n} R={r} T={t}/snapshot L={hist} B={2 * tq} queries/timestamp, "
                f"d={H_DIM}, 2-layer UnionRGCN + ConvTransE")
    return f"{args.workload} {args.model}"


_JSON_FD = None


def _quiet_stdout():
    """stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version there at WARN/VERSION
    level) are pointed at stderr for the whole run; emit() writes the line on the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    args = parse()
    _quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "all-entity-ranked queries/sec"

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
        cb = run_cpu_arm(args, steps, warmup)
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "queries/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(args), "variant": args.model, "note": "steps/warmup capped so the CPU arm "
                           "finishes in minutes; reference Python cannot travel to the GPU box, so the oracle port runs"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import regcn_b200 as R
    from regcn_b200 import _lib, evaluate, ops, synth, utils
    from regcn_b200 import dist as rdist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")                # NCCL prints its version on stdout at VERSION and WARN level
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    if args.gemm:
        ops.set_gemm_impl(args.gemm)

    cfg = model_cfg(args.model)
    case = synth.make_case(args.workload, rank)          # every rank owns different test timestamps (weak scaling)
    n, r = case["num_ents"], case["num_rels"]
    L, T = len(case["history"]), len(case["history"][0])
    model, sd = build_product_model(cfg, n, r, 0)
    model = model.to(dev)
    glist = [R.build_sub_graph(n, r, s, True, local) for s in case["history"]]
    test_dev = torch.from_numpy(case["test"]).to(dev)
    inv = test_dev[:, [2, 1, 0]].clone()
    inv[:, 1] += r
    all_t = torch.cat((test_dev, inv)).contiguous()
    B = all_t.shape[0]
    fcsr = utils.filter_csr_from_snapshot(all_t, 2 * r, 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, timers=False):
        for _ in range(warmup):
            fn(None)
        barrier()
        pairs, tms = [], []
        for _ in range(steps):
            flush.fill_(1.0)                               # L2 flush between timed iterations (untimed)
            tm = {k: (ev(), ev()) for k in ("evolve", "score", "rank")} if timers else None
            a, b = ev(), ev()
            a.record()
            fn(tm)
            b.record()
            pairs.append((a, b))
            tms.append(tm)
        barrier()
        tot = sum(a.elapsed_time(b) for a, b in pairs)
        parts = {}
        if timers:
            for k in ("evolve", "score", "rank"):
                parts[k] = sum(t[k][0].elapsed_time(t[k][1]) for t in tms) / steps
        return tot, parts

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    lib0 = _lib.load()
    l0 = None
    t_wall0 = time.time()
    tot_ms, parts = timed(lambda tm: evaluate.evaluate_snapshot(model, glist, all_t, fcsr, tm), args.steps,
                          args.warmup, timers=True)
    t_wall1 = time.time()
    # kernels launched by libregcn_b200.so in ONE device-resident step (counted inside the library's launcher)
    torch.cuda.synchronize()
    l0 = lib0.regcn_kernel_launches()
    evaluate.evaluate_snapshot(model, glist, all_t, fcsr, None)
    torch.cuda.synchronize()
    launches_per_step = int(lib0.regcn_kernel_launches() - l0)
    launches = launches_per_step * args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tot_ms = maxr(tot_ms)
    ms_per_step = tot_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    evolve_ms = maxr(parts["evolve"])

    # ---- end-to-end arm: host buffers -> public API -> host results --------------------------------------------
    # The public API is the reference's evaluation loop itself, regcn_b200.test() (src/main.py:33-123): a window of L
    # history snapshots slides over a stream of test snapshots held in PINNED HOST memory.  Every timed step copies
    # its test snapshot host->device, builds the edge index of the snapshot that entered the window, evolves, ranks
    # entities and relations (raw + time-filtered) and copies the four rank vectors device->host.
    e2e_steps = max(3, min(args.steps, 12))
    e2e_warm = max(L + 1, min(args.warmup, 3))       # the window must have turned over once (steady-state cache)
    stream = synth.make_stream(args.workload, 1000 + rank, n_test=e2e_warm + 2 * e2e_steps)
    s_hist = [torch.from_numpy(s).pin_memory() for s in stream["history"]]
    s_tests = [torch.from_numpy(s).pin_memory() for s in stream["tests"]]
    # warm-up: turn the window over, then one untimed call of exactly the timed call's shape (the pinned staging areas
    # and the caching allocators then hold blocks of the right sizes)
    R.test(model, s_hist, s_tests[:e2e_warm], r, n, True, test_history_len=L)
    R.test(model, (s_hist + s_tests[:e2e_warm])[-L:], s_tests[e2e_warm:e2e_warm + e2e_steps], r, n, True,
           test_history_len=L)
    e2e_warm += e2e_steps
    win = (s_hist + s_tests[:e2e_warm])[-L:]
    barrier()
    ea, eb = ev(), ev()
    ea.record()
    R.test(model, win, s_tests[e2e_warm:], r, n, True, test_history_len=L)
    eb.record()
    barrier()
    e2e_ms = maxr(ea.elapsed_time(eb)) / e2e_steps
    Bq = 2 * s_tests[0].shape[0]
    h2d = s_tests[0].numel() * 8
    d2h = 4 * Bq * 4 + 2 * 4 + 8 * 4

    # ---- roofline of the dominant kernel: the library records CUDA events around every launch of the tcgen05 GEMM
    #      (on the launching stream) while a few extra steps run; flops are the algorithmic 2*M*N*K of each launch ----
    import ctypes
    lib = _lib.load()
    probe_steps = 3
    torch.cuda.synchronize()
    # per-kernel timing brackets every launch with CUDA events on its stream: the probe steps run single-stream and
    # without programmatic dependent launch, so that a kernel's time is its own (the timed region above keeps both on)
    lib.regcn_two_stream_enable(0)
    lib.regcn_pdl_enable(0)
    lib.regcn_prof_enable(1)
    pa, pb = ev(), ev()
    probe_ms = 0.0
    for _ in range(probe_steps):
        flush.fill_(1.0)
        pa.record()
        evaluate.evaluate_snapshot(model, glist, all_t, fcsr, None)
        pb.record()
        torch.cuda.synchronize()
        probe_ms += pa.elapsed_time(pb)
    probe_ms /= probe_steps
    ms_c, n_c, w_c = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    lib.regcn_prof_read(0, ctypes.byref(ms_c), ctypes.byref(n_c), ctypes.byref(w_c))
    agg_ms_c, agg_n_c, agg_w_c = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    lib.regcn_prof_read(1, ctypes.byref(agg_ms_c), ctypes.byref(agg_n_c), ctypes.byref(agg_w_c))
    lib.regcn_prof_enable(0)
    lib.regcn_two_stream_enable(1)
    lib.regcn_pdl_enable(1)
    gemm_ms = ms_c.value / probe_steps
    gemm_flops = w_c.value / probe_steps
    n_gemm = n_c.value // probe_steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    traffic = None
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the GEMM kernel over one step, from the committed
        # `ncu --set full` capture of this same workload (profiles/README.md)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("gemm_tf32_kernel_bytes_per_launch")
    except Exception:
        pass
    ach_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    passes = 3 if ops.gemm_impl() == "tc" else 1
    roofline = {"kernel": ops.gemm_kernel_name(), "bound": "tensor", "achieved": ach_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks
                else "fallback 1.4 PFLOP/s (B200_PROFILING.md)",
                "launches_per_step": n_gemm, "ms_per_step_in_kernel": gemm_ms,
                "share_of_step": gemm_ms / probe_ms if probe_ms > 0 else None,
                "probe_step_ms": probe_ms,
                "executed_frac_of_tf32_peak": passes * ach_tf / (peak_tf / 2),
                "algorithmic_flops_per_step": gemm_flops,
                "note": f"average over the {n_gemm} GEMM launches of a step, most of them latency-bound at this size (N=23033 "
                        f"rows x 200..400 columns, relation GRU 512 rows); fp32-parity mode issues {passes} TF32 MMAs per "
                        f"algorithmic MAC, so executed tensor work is {passes}x the algorithmic flops and the TF32 dense peak is "
                        f"half the bf16 figure; probe steps run single-stream (see scoring_kernel for the kernel at a size "
                        f"where the tensor roofline binds)"}
    # edge kernel (HBM-bound) at this workload: algorithmic bytes per launch = 808*E + 808*N + 800*2R
    agg_launches = max(1, agg_n_c.value // probe_steps)
    agg_bytes = agg_launches * (808.0 * 2 * T + 808.0 * n + 800.0 * 2 * r)
    agg_ms = agg_ms_c.value / probe_steps
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    edge = {"kernel": "regcn::union_aggregate_kernel (+fixup)", "bound": "hbm", "launches_per_step": agg_launches,
            "ms_per_step_in_kernel": agg_ms, "achieved": agg_bytes / (agg_ms * 1e-3) / 1e9 if agg_ms > 0 else 0.0,
            "peak": hbm, "unit": "GB/s", "note": "latency-bound at this size (E=3082 edges, 18 MB output); see "
            "profiles/ for the HBM-bound stress sizes"}
    edge["frac"] = edge["achieved"] / hbm

    # ---- the two kernels north_star sets targets for, at sizes where their rooflines bind (rank 0, N = 1 only) ----
    stress = None
    if rank == 0 and world == 1 and not args.no_stress:
        stress = run_stress(dev, hbm, peak_tf)

    # ---- entity-sharded scoring + rank merge (strong scaling of one timestamp), all ranks on the same queries ----
    sharded = None
    if world > 1:
        case0 = synth.make_case(args.workload, 0)
        g0 = [R.build_sub_graph(n, r, s, True, local) for s in case0["history"]]
        t0 = torch.from_numpy(case0["test"]).to(dev)
        inv0 = t0[:, [2, 1, 0]].clone()
        inv0[:, 1] += r
        all0 = torch.cat((t0, inv0)).contiguous()
        f0 = utils.filter_csr_from_snapshot(all0, 2 * r, 0)
        embs, _, r_emb, _, _ = model.forward(g0, None, True)
        emb = ops.row_map(embs[-1], ops.ROW_NORMALIZE) if cfg["kind"] == "regcn" else embs[-1]
        out = {}

        def sharded_step(_):
            out["r"] = evaluate.score_rank_sharded(model, emb, r_emb, all0, f0)

        tot_s, _ = timed(sharded_step, args.steps, args.warmup)
        s_ms = maxr(tot_s) / args.steps
        rk1, frk1 = evaluate.evaluate_snapshot(model, g0, all0, f0, fused=True)
        same = bool(torch.equal(rk1, out["r"][0]) and torch.equal(frk1, out["r"][1]))
        sharded = {"queries_per_s": all0.shape[0] / (s_ms * 1e-3), "ms_per_step": s_ms, "scaling": "strong",
                   "ranks_equal_single_gpu": same, "collectives": "one all_reduce(SUM) of (2,B) int32 counts over NCCL",
                   "what": "query tower + fused score/count over N/G candidates per GPU + rank merge (evolution excluded)"}

    # ---- training step: get_loss in train() mode -> backward -> clip_grad_norm_(1.0) -> Adam, the
    #      reference's src/main.py:233-246 with its default dropout 0.2; every rank trains its own replica (no
    #      collective: the reference takes one optimiser step per snapshot, SURVEY 8e "replicas only") ----
    train_line = None
    if cfg["kind"] == "regcn":
        from regcn_b200 import optim as roptim
        tmodel, _ = build_product_model(cfg, n, r, 0)
        tmodel = tmodel.to(dev).train()
        topt = roptim.Adam(tmodel.parameters(), lr=1e-3, weight_decay=1e-5)
        out_t = {}

        def train_step(_):
            le, lr_, ls = tmodel.get_loss(glist, test_dev, None, True)
            (0.7 * le + 0.3 * lr_ + ls).backward()
            roptim.clip_grad_norm_(topt, 1.0)
            topt.step()
            topt.zero_grad()
            out_t["loss"] = le

        t_steps = max(3, min(args.steps, 10))
        tot_t, _ = timed(train_step, t_steps, 3)
        train_ms = maxr(tot_t) / t_steps
        torch.cuda.synchronize()
        l0 = lib0.regcn_kernel_launches()
        train_step(None)
        torch.cuda.synchronize()
        train_line = {"ms_per_step": train_ms, "train_snapshot_steps_per_s": world * L / (train_ms * 1e-3),
                      "optimisation_steps_per_s": world / (train_ms * 1e-3), "steps": t_steps,
                      "gpu_launches_per_step": int(lib0.regcn_kernel_launches() - l0),
                      "loss_ent_after": float(out_t["loss"].detach()),
                      "what": "RecurrentRGCN.get_loss (train mode, dropout 0.2, batch-stat BatchNorm) + backward + "
                              "clip_grad_norm_(1.0) + Adam(lr 1e-3, wd 1e-5); 3xTF32 tcgen05 GEMMs for forward, dX and dW"}
        del tmodel, topt

    # ---- entity-sharded scoring at BASELINE configs[4] size (1M entities, 8192 queries, RotH-form hyperbolic score):
    #      each rank counts over its N/G candidate rows with the fused kernel, ONE all_reduce(SUM) of the (B,) counts;
    #      rank 0 also times the whole table on one GPU, so the line carries its own strong-scaling figure ----
    sharded_c5 = None
    if world > 1:
        Bq5, N5, d5 = 8192, 1_000_000, H_DIM
        lo5, hi5 = rdist.shard_bounds(N5, rank, world)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234)
        q5 = torch.randn(Bq5, d5, device=dev, generator=gen) * 0.05
        gen.manual_seed(99 + rank)
        e5 = torch.randn(hi5 - lo5, d5, device=dev, generator=gen) * 0.05
        qh5, ql5 = ops.split_tf32(q5)
        eh5, el5 = ops.split_tf32(e5)
        x2_5, y2_5 = ops.row_sumsq(q5), ops.row_sumsq(e5)
        sm5 = torch.tensor([1.0, 1.0], device=dev)
        tgt5 = torch.randint(0, N5, (Bq5,), device=dev, dtype=torch.int32, generator=gen)
        ts5 = torch.zeros(Bq5, device=dev)
        raw5 = torch.zeros(Bq5, device=dev, dtype=torch.int32)

        def sharded5(_):
            raw5.zero_()
            _lib.call("regcn_score_count_tf32", qh5.data_ptr(), ql5.data_ptr(), eh5.data_ptr(), el5.data_ptr(), Bq5,
                      hi5 - lo5, d5, ts5.data_ptr(), tgt5.data_ptr(), raw5.data_ptr(), lo5, 1, x2_5.data_ptr(),
                      y2_5.data_ptr(), None, 0.01, sm5.data_ptr(), None, 3)
            dist.all_reduce(raw5)

        n5 = max(3, min(args.steps, 10))
        tot5, _ = timed(sharded5, n5, 3)
        ms5 = maxr(tot5) / n5
        single_ms = None
        if rank == 0:
            gen.manual_seed(7)
            eF = torch.randn(N5, d5, device=dev, generator=gen) * 0.05
            eFh, eFl = ops.split_tf32(eF)
            y2F = ops.row_sumsq(eF)
            del eF

            def full5():
                raw5.zero_()
                _lib.call("regcn_score_count_tf32", qh5.data_ptr(), ql5.data_ptr(), eFh.data_ptr(), eFl.data_ptr(), Bq5,
                          N5, d5, ts5.data_ptr(), tgt5.data_ptr(), raw5.data_ptr(), 0, 1, x2_5.data_ptr(), y2F.data_ptr(),
                          None, 0.01, sm5.data_ptr(), None, 3)

            for _ in range(2):
                full5()
            torch.cuda.synchronize()
            fa, fb = ev(), ev()
            fa.record()
            for _ in range(3):
                full5()
            fb.record()
            torch.cuda.synchronize()
            single_ms = fa.elapsed_time(fb) / 3
            del eFh, eFl, y2F
        barrier()
        sharded_c5 = {"shape": f"B={Bq5} queries x N={N5} entities, d={d5}, hyperbolic (RotH-form) score, 3xTF32",
                      "ms_per_step": ms5, "queries_per_s": Bq5 / (ms5 * 1e-3), "single_gpu_full_table_ms": single_ms,
                      "speedup_vs_one_gpu": (single_ms / ms5) if single_ms else None,
                      "strong_scaling_efficiency": (single_ms / ms5 / world) if single_ms else None,
                      "collectives": "one all_reduce(SUM) of (B,) int32 counts over NCCL"}
        del q5, e5, qh5, ql5, eh5, el5

    # ---- the reference's only published number (hyperbolic_src/train.log: hyperbolic_uvrgcn + hyperbolic_convtranse,
    #      ICEWS14s, layer_norm, history 3: 61.9-69.2 s per epoch of 303 snapshot steps = 4.9 optimisation steps/s on an
    #      unnamed GPU): the same configuration on the ICEWS14s-shaped synthetic workload c1 ----
    train_hyp_line = None
    if rank == 0 and world == 1 and not args.no_stress:
        from regcn_b200 import optim as roptim
        hcase = synth.make_case("c1", 0)
        hn, hr = hcase["num_ents"], hcase["num_rels"]
        hm = R.HyperbolicRecurrentRGCN("hyperbolic_convtranse", "hyperbolic_uvrgcn", hn, hr, 0, 0, H_DIM, "sub", 3,
                                       num_bases=N_BASES, num_hidden_layers=N_LAYERS, dropout=0.2, c=0.01, self_loop=True,
                                       layer_norm=True, input_dropout=0.2, hidden_dropout=0.2, feat_dropout=0.2,
                                       entity_prediction=True, relation_prediction=True, use_cuda=True, gpu=0)
        hm.load_state_dict(synth.fill_state_dict(hm.state_dict(), 0))
        hm = hm.to(dev).train()
        hopt = roptim.Adam(hm.parameters(), lr=1e-3, weight_decay=1e-5)
        hg = [R.build_sub_graph(hn, hr, s, True, local) for s in hcase["history"]]
        ht = torch.from_numpy(hcase["test"]).to(dev)

        def hyp_train_step(_):
            le, lr_, ls, lrad = hm.get_loss(hg, ht, None, True)
            (0.7 * le + 0.3 * lr_ + ls + lrad).backward()
            roptim.clip_grad_norm_(hopt, 1.0)
            hopt.step()
            hopt.zero_grad()

        h_steps = max(3, min(args.steps, 10))
        tot_h, _ = timed(hyp_train_step, h_steps, 3)
        h_ms = tot_h / h_steps
        train_hyp_line = {"workload": "c1 (ICEWS14s-shaped: N=7128 R=230 T=250/snapshot, history 3), hyperbolic_uvrgcn + "
                                      "hyperbolic_convtranse, layer_norm, dropout 0.2, whole snapshot per step",
                          "ms_per_step": h_ms, "optimisation_steps_per_s": 1e3 / h_ms,
                          "train_snapshot_steps_per_s": len(hg) * 1e3 / h_ms,
                          "reference_published_steps_per_s": 4.9,
                          "reference_source": "hyperbolic_src/train.log:36-44 (unnamed GPU, real ICEWS14s, fwd+bwd+Adam)",
                          "ratio_to_published": (1e3 / h_ms) / 4.9}
        del hm, hopt

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = run_cpu_arm(args, 3, 1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if train_line is not None:
            ct = run_cpu_train_arm(args, 1, 1)
            train_line["cpu_baseline"] = ct

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(args),
                           "variant": args.model, "parallelism": f"timestamp-dp{world}",
                           "l2": "256 MiB buffer written between timed steps (untimed)",
                           "gemm_impl": ops.gemm_impl()},
                "snapshot_steps_per_s": world * L / (evolve_ms * 1e-3), "evolve_ms_per_step": evolve_ms,
                "phase_ms": parts, "e2e": {"value": world * Bq / (e2e_ms * 1e-3), "unit": "queries/s",
                                           "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                                           "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                                           "api": "regcn_b200.test(): sliding-window loop of src/main.py:33-123 over "
                                                  "pinned host snapshots (entity + relation ranks, raw + filtered); "
                                                  "per-step working set ~480 MB > L2"},
                "gpu_launches": launches, "roofline": roofline, "edge_kernel": edge, "clocks": clocks}
        line["gpu_launches_per_step"] = launches_per_step
        if stress:
            line.update(stress)
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if sharded:
            line["entity_sharded"] = sharded
        if sharded_c5:
            line["entity_sharded_c5"] = sharded_c5
        if train_line:
            line["train"] = train_line
        if train_hyp_line:
            line["train_hyperbolic"] = train_hyp_line
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
