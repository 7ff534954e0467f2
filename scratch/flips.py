import sys, numpy as np, torch
sys.path.insert(0,'.')
from oracle import restate, synth
from tests.helpers import build_model, load_golden
import regcn_b200 as R
from regcn_b200 import ops, utils
for name in ["regcn_c3_s1","regcn_c1_s0","hyp_lgcn_roth_c1_s0"]:
    cfg, z = load_golden(name)
    case = synth.make_case(cfg["shape"], cfg["seed"]); n, r = case["num_ents"], case["num_rels"]
    for impl in ["simt","tc"]:
        ops.set_gemm_impl(impl)
        model, sd = build_model(cfg, n, r); model = model.cuda()
        gl = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
        all_t, score, score_rel = model.predict(gl, r, None, torch.from_numpy(case["test"]).cuda(), True)
        all_ans = synth.answers_of(case["test"], r, False)
        fm, m, rank, frank = utils.get_total_rank(all_t, score.clone(), all_ans, 1000, 0)
        d = (rank.cpu().numpy() - z["rank"]); 
        blk = score[torch.as_tensor(z["sub_qrows"]).cuda()][:, torch.as_tensor(z["sub_rows"]).cuda()].cpu().numpy()
        print(name, impl, "flips %.3f%%" % (100*np.mean(d!=0)), "max|drank|", np.abs(d).max(), "mean|drank|", np.abs(d).mean(),
              "mrr mine %.6f ref %.6f" % (m, z["mrr"][1]), "score blk maxabs err %.2e" % np.abs(blk - z["score_block"]).max(), "score scale %.2f" % np.abs(z["score_block"]).max())
