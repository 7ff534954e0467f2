"""Device time of the ConvTransE query tower at the ICEWS18 shape: fused (regcn_convtrans_fc) vs feature map through memory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import regcn_b200 as R
from regcn_b200 import _lib, synth, ops
from bench import build_product_model, model_cfg
_lib.require_device()
dev = torch.device("cuda", 0)
case = synth.make_case("c3", 0)
n, r = case["num_ents"], case["num_rels"]
m, _ = build_product_model(model_cfg("regcn"), n, r, 0)
m = m.to(dev)
t = torch.from_numpy(case["test"]).to(dev)
inv = t[:, [2, 1, 0]].clone(); inv[:, 1] += r
all_t = torch.cat((t, inv)).contiguous()
emb = torch.randn(n, 200, device=dev); rel = torch.randn(2 * r, 200, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("1", "0"):
    os.environ["REGCN_FUSED_TOWER"] = mode
    ts = []
    for rep in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        q = m.decoder_ob._tower(emb, rel, all_t, 0, 1, always_bn2=False)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000)
    print("fused" if mode == "1" else "feature map", [round(x, 1) for x in sorted(ts[2:])])
