"""In-kernel timeline of the tcgen05 GEMM at the shapes of one C3 evolve step (regcn_gemm_tf32_trace): where does a
40-55 us layer GEMM spend its time -- dependency wait, operand loads, MMA issue, epilogue?  Each shape runs warm
(L2-resident operands, like snapshot 2..L of a step) and is timed with CUDA events; the last launch is traced.

    python profiles/gemm_trace.py            (writes gpurun_out/gemm_trace.json)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from regcn_b200 import _lib, ops

lib = _lib.load()
_lib.require_device()
dev = torch.device("cuda", 0)
SLOTS = lib.regcn_gemm_tf32_trace_slots()
N, d, R2 = int(os.environ.get("REGCN_TRACE_ROWS", "23033")), 200, 512      # REGCN_TRACE_ROWS=184264: 8 windows per recurrence
n_act = 1560 * max(1, N // 23033)
rng = torch.Generator(device=dev)
rng.manual_seed(0)


def rnd(*s):
    return torch.randn(*s, device=dev, generator=rng) * 0.1


def split(x):
    return ops.split_tf32(x.contiguous())


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def trace(fn):
    buf = torch.zeros(148 * SLOTS, device=dev, dtype=torch.int64)
    torch.cuda.synchronize()
    lib.regcn_gemm_tf32_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.regcn_gemm_tf32_trace(None)
    t = buf.cpu().numpy().reshape(148, SLOTS)
    used = t[:, 0] > 0
    t = t[used].astype(np.float64)
    t0 = t[:, 0].min()
    rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
    return rel


def describe(name, rel, us):
    g = rel.shape[0]
    med = lambda c: float(np.nanmedian(rel[:, c])) if np.isfinite(rel[:, c]).any() else None
    mx = lambda c: float(np.nanmax(rel[:, c])) if np.isfinite(rel[:, c]).any() else None
    out = {"name": name, "event_us": us, "grid": g, "entry_med": med(0), "after_wait_med": med(1), "end_med": med(40),
           "end_max": mx(40), "tiles": []}
    for it in range(6):
        if not np.isfinite(rel[:, 11 + 3 * it]).any():
            break
        out["tiles"].append({"load_start": med(2 + it) if it < 4 else None, "load_issued": med(6 + it) if it < 4 else None,
                             "mma_slot_free": med(10 + 3 * it), "first_operands": med(11 + 3 * it),
                             "mma_issued": med(12 + 3 * it), "acc_complete": med(28 + 2 * it),
                             "epi_done": med(29 + 2 * it)})
    print(f"== {name}: {us:.1f} us/launch (events, back to back), grid {g}; medians over CTAs, us from first CTA entry")
    print(f"   entry {out['entry_med']:.2f}  after-wait {out['after_wait_med']:.2f}  end med {out['end_med']:.2f} max {out['end_max']:.2f}")
    for i, t in enumerate(out["tiles"]):
        f = lambda v: "  -  " if v is None else f"{v:6.2f}"
        print(f"   tile {i}: loads {f(t['load_start'])}..{f(t['load_issued'])}  mma: slot-free {f(t['mma_slot_free'])} first-operands "
              f"{f(t['first_operands'])} issued {f(t['mma_issued'])}  acc-complete {f(t['acc_complete'])}  epilogue-done {f(t['epi_done'])}")
    return out


def main():
    res = []
    x = rnd(N, d)
    xh, xl = split(x)
    w_ev_t = rnd(2 * d, d)           # [W_evolve | W_time]^T  (N, K) layout
    wh, wl = split(w_ev_t)
    w1h, w1l = split(rnd(d, d))
    skip = torch.full((N,), -1, device=dev, dtype=torch.int32)
    act_rows = torch.randperm(N, device=dev)[:n_act].sort().values.to(torch.int32)
    skip[act_rows.long()] = torch.arange(n_act, device=dev, dtype=torch.int32)
    o_hi, o_lo, o_raw = torch.empty(N, d, device=dev), torch.empty(N, d, device=dev), torch.empty(N, d, device=dev)
    gate = torch.empty(N, d, device=dev)
    gbias = rnd(d)
    hprev = rnd(N, d)

    def le0():
        _lib.call("regcn_gemm_tf32_layer", xh.data_ptr(), xl.data_ptr(), d, wh.data_ptr(), wl.data_ptr(), d, N, 2 * d, d, d,
                  None, o_hi.data_ptr(), o_lo.data_ptr(), gate.data_ptr(), d, None, skip.data_ptr(), None, 0, None, None, 0)

    def le1():
        _lib.call("regcn_gemm_tf32_layer", xh.data_ptr(), xl.data_ptr(), d, w1h.data_ptr(), w1l.data_ptr(), d, N, d, d, d,
                  o_raw.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr(), None, 0, None, skip.data_ptr(), gate.data_ptr(), d,
                  gbias.data_ptr(), hprev.data_ptr(), 1)

    C2 = torch.empty(N, 2 * d, device=dev)
    C1 = torch.empty(N, d, device=dev)

    def store400():
        _lib.call("regcn_gemm_tf32", xh.data_ptr(), xl.data_ptr(), d, wh.data_ptr(), wl.data_ptr(), d, C2.data_ptr(), 2 * d,
                  N, 2 * d, d, None, 0, 3, 1, None, 0)

    def store200():
        _lib.call("regcn_gemm_tf32", xh.data_ptr(), xl.data_ptr(), d, w1h.data_ptr(), w1l.data_ptr(), d, C1.data_ptr(), d, N, d,
                  d, None, 0, 3, 1, None, 0)

    def store200_1pass():
        _lib.call("regcn_gemm_tf32", xh.data_ptr(), None, d, w1h.data_ptr(), None, d, C1.data_ptr(), d, N, d, d, None, 0, 1, 1,
                  None, 0)

    xm_h, xm_l = split(rnd(R2, d))
    wih_h, wih_l = split(rnd(3 * d, d))
    gi = torch.empty(R2, 3 * d, device=dev)
    bias3 = rnd(3 * d)

    def gru_gh():
        _lib.call("regcn_gemm_tf32", xm_h.data_ptr(), xm_l.data_ptr(), d, wih_h.data_ptr(), wih_l.data_ptr(), d, gi.data_ptr(),
                  3 * d, R2, 3 * d, d, bias3.data_ptr(), 0, 3, 1, None, 0)

    ag_h, ag_l = split(rnd(n_act, 2 * d))
    wc_h, wc_l = split(rnd(d, 2 * d))

    def compact0():
        _lib.call("regcn_gemm_tf32_layer", ag_h.data_ptr(), ag_l.data_ptr(), 2 * d, wc_h.data_ptr(), wc_l.data_ptr(), 2 * d,
                  n_act, d, 2 * d, d, o_raw.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr(), None, 0, act_rows.data_ptr(), None,
                  None, 0, None, None, 0)

    P = torch.empty(n_act, d, device=dev)

    def compact1():
        _lib.call("regcn_gemm_tf32", ag_h.data_ptr(), ag_l.data_ptr(), 2 * d, wc_h.data_ptr(), wc_l.data_ptr(), 2 * d,
                  P.data_ptr(), d, n_act, d, 2 * d, None, 0, 3, 1, None, 0)

    def le0_a32():
        _lib.call("regcn_gemm_tf32_layer_a32", x.data_ptr(), d, d, None, None, 0, 0, None, wh.data_ptr(), wl.data_ptr(), d, N,
                  2 * d, d, o_raw.data_ptr(), None, None, gate.data_ptr(), d, None, skip.data_ptr(), None, 0, None, None, 0)

    def le1_a32():
        _lib.call("regcn_gemm_tf32_layer_a32", x.data_ptr(), d, d, None, None, 0, 0, None, w1h.data_ptr(), w1l.data_ptr(), d, N,
                  d, d, o_raw.data_ptr(), None, None, None, 0, None, skip.data_ptr(), gate.data_ptr(), d, gbias.data_ptr(),
                  hprev.data_ptr(), 1)

    def store200_a32():
        _lib.call("regcn_gemm_tf32_a32", x.data_ptr(), d, d, None, None, 0, 0, None, w1h.data_ptr(), w1l.data_ptr(), d,
                  C1.data_ptr(), d, N, d, None, 0, 3, 1, None, 0, None, 0)

    agg = rnd(n_act, d)

    def compact0_a32():
        _lib.call("regcn_gemm_tf32_layer_a32", agg.data_ptr(), d, d, None, x.data_ptr(), d, d, act_rows.data_ptr(),
                  wc_h.data_ptr(), wc_l.data_ptr(), 2 * d, n_act, d, d, o_raw.data_ptr(), None, None, None, 0,
                  act_rows.data_ptr(), None, None, 0, None, None, 0)

    cases = [("A32 Le0 23033x400x200 (fp32 A, raw out)", le0_a32), ("A32 Le1 gate 23033x200x200 (fp32 A, raw out)", le1_a32),
             ("A32 store 23033x200x200", store200_a32), ("A32 compact0 [agg|x gathered] 1560x200x400", compact0_a32),
             ("Le0 layer-epilogue 23033x400x200", le0), ("Le1 gate-epilogue 23033x200x200", le1),
             ("store 23033x400x200", store400), ("store 23033x200x200", store200),
             ("store 23033x200x200 1-pass", store200_1pass), ("GRU gh 512x600x200 +bias", gru_gh),
             ("compact0 1560x200x400 scatter-epilogue", compact0), ("compact1 1560x200x400 store", compact1)]
    cap = int(os.environ.get("REGCN_TRACE_CAP", "0"))          # persistent-grid cap: per-SM vs chip-wide limits
    only = os.environ.get("REGCN_TRACE_ONLY")
    lib.regcn_gemm_tf32_grid_cap(cap)
    for name, fn in cases:
        if only and only not in name:
            continue
        us = timed(fn)
        res.append(describe(name + (f" [grid cap {cap}]" if cap else ""), trace(fn), us))
    lib.regcn_gemm_tf32_grid_cap(0)
    if only:
        return
    # tile-shape sensitivity of the two big layer GEMMs
    for bn, st in ((208, 0), (128, 0), (104, 0), (64, 0)):
        lib.regcn_gemm_tf32_tune(bn, st)
        try:
            res.append({"name": f"store 23033x200x200 block_n={bn}", "event_us": timed(store200)})
            res.append({"name": f"Le0 block_n={bn}", "event_us": timed(le0)})
            print(res[-2], res[-1])
        except RuntimeError as e:
            print("tune", bn, "failed:", e)
    lib.regcn_gemm_tf32_tune(0, 0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_trace.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
