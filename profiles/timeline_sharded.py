"""Kernel timeline (rank 0) of ONE entity-sharded scoring step of a C3 timestamp under torchrun:
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 profiles/timeline_sharded.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
import regcn_b200 as R
from regcn_b200 import _lib, evaluate, ops, synth, utils
from bench import build_product_model, model_cfg
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
case = synth.make_case("c3", 0); n, r = case["num_ents"], case["num_rels"]
model, _ = build_product_model(model_cfg("regcn"), n, r, 0); model = model.to(dev)
g0 = [R.build_sub_graph(n, r, s, True, local) for s in case["history"]]
t0 = torch.from_numpy(case["test"]).to(dev); inv = t0[:, [2, 1, 0]].clone(); inv[:, 1] += r
all0 = torch.cat((t0, inv)).contiguous(); f0 = utils.filter_csr_from_snapshot(all0, 2 * r, 0)
embs, _, r_emb, _, _ = model.forward(g0, None, True)
emb = ops.row_map(embs[-1], ops.ROW_NORMALIZE)
for _ in range(5):
    evaluate.score_rank_sharded(model, emb, r_emb, all0, f0)
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        evaluate.score_rank_sharded(model, emb, r_emb, all0, f0)
        torch.cuda.synchronize()
if dist.get_rank() == 0:
    path = "gpurun_out/trace_sharded.json"; os.makedirs("gpurun_out", exist_ok=True)
    prof.export_chrome_trace(path)
    ev = sorted((e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"])
    os.remove(path)
    last = ev[-len(ev) // 3:]
    t_0 = last[0]["ts"]
    for e in last:
        print(f"t={e['ts'] - t_0:8.1f} dur={e['dur']:7.1f} s={e['args'].get('stream')} {e['name'][:90]}")
dist.destroy_process_group()
