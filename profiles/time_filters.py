"""Device time of the time-aware filter-list kernels at the ICEWS18 shape (one test snapshot incl. inverses):
regcn_queries_prepare (inverse triples + both count passes + scans) and the two fills; list-length histogram."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from regcn_b200 import _lib, synth, utils
_lib.require_device()
dev = torch.device("cuda", 0)
stream = synth.make_stream("c3", 1000, n_test=12)
r = stream["num_rels"]
tests = [torch.from_numpy(s).to(dev) for s in stream["tests"]]


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000)
    return sorted(ts)[len(ts) // 2], out


for t in tests[:2]:
    us, (all_t, pe, pr, totals) = timed(lambda: utils.queries_prepare(t, r))
    tot = totals.tolist()
    us_e, fe = timed(lambda: pe.finish(tot[0]))
    us_r, fr = timed(lambda: pr.finish(tot[1]))
    # back to back (how test() enqueues them)
    def both():
        for _ in range(12):
            utils.filter_lists_finish2(pe, tot[0], pr, tot[1])
    us_b, _ = timed(both, 5)
    le = (fe.end - fe.ptr).cpu(); lr = (fr.end - fr.ptr).cpu()
    ce = (pe.beg[1:] - pe.beg[:-1]).cpu(); cr = (pr.beg[1:] - pr.beg[:-1]).cpu()
    print(f"B={all_t.shape[0]} prepare {us:.1f} us, fill ent {us_e:.1f} us, fill rel {us_r:.1f} us, 12 x regcn_filter_fill2 back to back {us_b / 12:.1f} us per timestamp")
    print("  matches per query (ent) max", int(ce.max()), "mean %.2f" % float(ce.float().mean()), " >32:", int((ce > 32).sum()), " >256:", int((ce > 256).sum()),
          "| (rel) max", int(cr.max()), "mean %.2f" % float(cr.float().mean()), " >32:", int((cr > 32).sum()), " >256:", int((cr > 256).sum()))
    print("  unique per query (ent) max", int(le.max()), "(rel) max", int(lr.max()))
