"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py            # regenerate every fixture

The reference modules (src/rrgcn.py, rgcn/{layers,utils}.py, src/decoder.py, hyperbolic_src/*) are imported as
they lie under /root/reference with the DGL stand-in of oracle/fake_dgl.py; inputs and parameters come from
oracle/synth.py (numpy default_rng, reproducible anywhere), so a fixture stores only its config, seed and the
reference's outputs.  Large cases store a seeded subsample of rows / columns plus the full integer outputs.
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REGCN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import fake_dgl, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# fixture name -> config
CASES = {
    "regcn_tiny_s0": dict(kind="regcn", shape="tiny", seed=0, layer_norm=True),
    "regcn_tiny_s1_noln": dict(kind="regcn", shape="tiny", seed=1, layer_norm=False),
    "regcn_small_s2": dict(kind="regcn", shape="small", seed=2, layer_norm=True),
    "regcn_c1_s0": dict(kind="regcn", shape="c1", seed=0, layer_norm=True, sub=96),
    "regcn_c3_s1": dict(kind="regcn", shape="c3", seed=1, layer_norm=True, sub=64),
    "hyp_uv_convtranse_tiny_s0": dict(kind="hyp", shape="tiny", seed=0, encoder="hyperbolic_uvrgcn",
                                      decoder="hyperbolic_convtranse", layer_norm=False, gamma=0.15),
    "hyp_uv_convtranse_tiny_s1_ln": dict(kind="hyp", shape="tiny", seed=1, encoder="hyperbolic_uvrgcn",
                                         decoder="hyperbolic_convtranse", layer_norm=True, gamma=1.0),
    "hyp_uv_murp_tiny_s2": dict(kind="hyp", shape="tiny", seed=2, encoder="hyperbolic_uvrgcn", decoder="murp",
                                layer_norm=False, gamma=0.15),
    "hyp_lgcn_roth_tiny_s0": dict(kind="hyp", shape="tiny_l", seed=0, encoder="lgcn", decoder="roth",
                                  layer_norm=False, gamma=0.15),
    "hyp_lgcn_roth_small_s1": dict(kind="hyp", shape="small_l", seed=1, encoder="lgcn", decoder="roth",
                                   layer_norm=False, gamma=0.15),
    "hyp_lgcn_roth_c1_s0": dict(kind="hyp", shape="c1", seed=0, encoder="lgcn", decoder="roth", layer_norm=False,
                                gamma=0.15, sub=96),
    "hyp_uv_roth_tiny_flags_s3": dict(kind="hyp", shape="tiny", seed=3, encoder="hyperbolic_uvrgcn", decoder="roth",
                                      layer_norm=False, gamma=0.15, entity_bias=True, rel_curvature=True),
    "hyp_uv_murp_tiny_flags_s4": dict(kind="hyp", shape="tiny", seed=4, encoder="hyperbolic_uvrgcn", decoder="murp",
                                      layer_norm=True, gamma=0.15, entity_bias=True, rel_curvature=True),
    "hyp_lgcn_roth_tiny_bias_s5": dict(kind="hyp", shape="tiny_l", seed=5, encoder="lgcn", decoder="roth",
                                       layer_norm=False, gamma=0.15, entity_bias=True),
    "hyp_uv_atth_tiny_s6": dict(kind="hyp", shape="tiny", seed=6, encoder="hyperbolic_uvrgcn", decoder="atth",
                                layer_norm=False, gamma=0.15),
    "hyp_lgcn_atth_small_flags_s7": dict(kind="hyp", shape="small_l", seed=7, encoder="lgcn", decoder="atth",
                                         layer_norm=False, gamma=0.15, entity_bias=True, rel_curvature=True),
    "hyp_uv_roth_c1_s1": dict(kind="hyp", shape="c1", seed=1, encoder="hyperbolic_uvrgcn", decoder="roth",
                              layer_norm=True, gamma=0.15, sub=96),
    # GDELT shape with dense snapshots (BASELINE configs[3]: 5000 triples per snapshot, hub rows of ~900 in-edges)
    "regcn_c4_s0": dict(kind="regcn", shape="c4", seed=0, layer_norm=True, sub=96),
    # the hyperbolic model of BASELINE configs[1] at the ICEWS18 shape of configs[2]
    "hyp_lgcn_roth_c3_s2": dict(kind="hyp", shape="c3", seed=2, encoder="lgcn", decoder="roth", layer_norm=False,
                                gamma=0.15, sub=64),
}
H_DIM = 200
N_BASES = 100
N_LAYERS = 2
CURV = 0.01


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not present: golden fixtures can only be regenerated in the build container")
    fake_dgl.install()
    sys.path.insert(0, REF)
    import logging
    logging.disable(logging.CRITICAL)
    from rgcn import utils as ref_utils
    from src.rrgcn import RecurrentRGCN
    from hyperbolic_src.hyperbolic_model import HyperbolicRecurrentRGCN
    return ref_utils, RecurrentRGCN, HyperbolicRecurrentRGCN


def build_reference_model(cfg, n, r, RecurrentRGCN, HyperbolicRecurrentRGCN):
    if cfg["kind"] == "regcn":
        m = RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, H_DIM, "sub", 3, num_bases=N_BASES, num_basis=-1,
                          num_hidden_layers=N_LAYERS, dropout=0.2, self_loop=True, skip_connect=False,
                          layer_norm=cfg["layer_norm"], input_dropout=0.2, hidden_dropout=0.2, feat_dropout=0.2,
                          entity_prediction=True, relation_prediction=True, use_cuda=False, gpu="cpu")
    else:
        m = HyperbolicRecurrentRGCN(cfg["decoder"], cfg["encoder"], n, r, 0, 0, H_DIM, "sub", 3, num_bases=N_BASES,
                                    num_hidden_layers=N_LAYERS, dropout=0.2, c=CURV, self_loop=True,
                                    skip_connect=False, layer_norm=cfg["layer_norm"], input_dropout=0.2,
                                    hidden_dropout=0.2, feat_dropout=0.2, entity_prediction=True,
                                    relation_prediction=True, use_cuda=False, gpu="cpu",
                                    radius_msg_gamma=cfg["gamma"], hyp_init_scale=1e-3,
                                    use_entity_euclidean_bias=cfg.get("entity_bias", False),
                                    use_relation_specific_curvature=cfg.get("rel_curvature", False))
    m.load_state_dict(synth.fill_state_dict(m.state_dict(), cfg["seed"]))
    m.eval()
    return m


def run_case(name, cfg, ref_utils, RecurrentRGCN, HyperbolicRecurrentRGCN):
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    model = build_reference_model(cfg, n, r, RecurrentRGCN, HyperbolicRecurrentRGCN)
    glist = [ref_utils.build_sub_graph(n, r, snap, False, "cpu") for snap in case["history"]]
    out = {"config": np.array(json.dumps(cfg))}
    # ---- integer outputs of build_sub_graph / r2e (rgcn/utils.py:78-134) for every snapshot
    for i, g in enumerate(glist):
        src, dst = g.edges()
        out[f"g{i}_src"] = src.numpy().astype(np.int32)
        out[f"g{i}_dst"] = dst.numpy().astype(np.int32)
        out[f"g{i}_type"] = g.edata["type"].numpy().astype(np.int32)
        out[f"g{i}_norm"] = g.ndata["norm"].view(-1).numpy().astype(np.float32)
        out[f"g{i}_uniq_r"] = np.asarray(g.uniq_r, dtype=np.int32)
        out[f"g{i}_r_len"] = np.asarray(g.r_len, dtype=np.int32).reshape(-1, 2)
        out[f"g{i}_r_to_e"] = np.asarray([int(x) for x in g.r_to_e], dtype=np.int32)
    test = torch.from_numpy(case["test"])
    t0 = time.time()
    with torch.no_grad():
        all_triples, score, score_rel = model.predict(glist, r, None, test, False)
        hist, _, h0, _, _ = model.forward(glist, None, False)
    dt = time.time() - t0
    all_ans = ref_utils.load_all_answers_for_filter(case["test"], r, False)
    all_ans_r = ref_utils.load_all_answers_for_filter(case["test"], r, True)
    fm, m, rank, frank = ref_utils.get_total_rank(all_triples, score.clone(), all_ans, eval_bz=1000, rel_predict=0)
    fmr, mr, rank_r, frank_r = ref_utils.get_total_rank(all_triples, score_rel.clone(), all_ans_r, eval_bz=1000,
                                                          rel_predict=1)
    out.update(all_triples=all_triples.numpy().astype(np.int64), rank=rank.numpy(), filter_rank=frank.numpy(),
               rank_rel=rank_r.numpy(), filter_rank_rel=frank_r.numpy(),
               mrr=np.array([fm, m, fmr, mr], dtype=np.float64), h0=h0.numpy())
    sub = cfg.get("sub")
    if sub:
        rng = np.random.default_rng(777 + cfg["seed"])
        rows = np.sort(rng.choice(n, size=sub, replace=False))
        qrows = np.sort(rng.choice(score.shape[0], size=min(sub, score.shape[0]), replace=False))
        out.update(sub_rows=rows, sub_qrows=qrows, hist_last_rows=hist[-1][rows].numpy(),
                   hist_first_rows=hist[0][rows].numpy(), score_block=score[qrows][:, rows].numpy(),
                   score_rel_qrows=score_rel[qrows].numpy(),
                   score_rowsum=score.double().sum(1).numpy(), score_absmax=score.abs().max(1).values.numpy())
        # complete score rows (every candidate) of 16 seeded queries: the gate on scores covers whole rows, not a block
        frows = np.sort(rng.choice(score.shape[0], size=min(16, score.shape[0]), replace=False))
        out.update(full_qrows=frows, score_full_rows=score[frows].numpy())
    else:
        out.update(hist=np.stack([h.numpy() for h in hist]), score=score.numpy(), score_rel=score_rel.numpy())
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: reference ran in {dt:.2f}s -> {path} ({os.path.getsize(path) / 1e6:.2f} MB), "
          f"mrr(filter/raw)={fm:.4f}/{m:.4f}")


def run_losses(ref_utils, RecurrentRGCN, HyperbolicRecurrentRGCN):
    """tests/golden/losses.json: the reference's get_loss() (src/rrgcn.py:197-248, hyperbolic_model.py:941-1088) on
    every golden case, eval mode (dropout off, rrelu at its eval slope) under no_grad, with the test snapshot as the
    training triples.  Values: [loss_ent, loss_rel, loss_static(, loss_radius)]."""
    out = {}
    for name, cfg in CASES.items():
        case = synth.make_case(cfg["shape"], cfg["seed"])
        n, r = case["num_ents"], case["num_rels"]
        model = build_reference_model(cfg, n, r, RecurrentRGCN, HyperbolicRecurrentRGCN)
        glist = [ref_utils.build_sub_graph(n, r, snap, False, "cpu") for snap in case["history"]]
        with torch.no_grad():
            losses = model.get_loss(glist, torch.from_numpy(case["test"]), None, False)
        out[name] = [float(x.reshape(-1)[0]) for x in losses]
        print(name, out[name])
    with open(os.path.join(GOLDEN, "losses.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


# ---------------------------------------------------------------------------------------------------------------
# training golden: the reference's optimisation step (src/main.py:235-246) executed here, dropout 0
# ---------------------------------------------------------------------------------------------------------------
TRAIN_CASES = {
    "regcn_tiny_s0": dict(kind="regcn", shape="tiny", seed=0, layer_norm=True),
    "regcn_tiny_s1_noln": dict(kind="regcn", shape="tiny", seed=1, layer_norm=False),
    "regcn_small_s2": dict(kind="regcn", shape="small", seed=2, layer_norm=True),
    # --skip-connect: the uvrgcn cell calls every layer with prev_h=[] (src/rrgcn.py:37-38), so the gate weights exist,
    # are never used and never receive a gradient
    "regcn_tiny_s3_skip": dict(kind="regcn", shape="tiny", seed=3, layer_norm=True, skip_connect=True),
}
TRAIN_STEPS = 2
TASK_WEIGHT, GRAD_NORM, LR, WEIGHT_DECAY = 0.7, 1.0, 1e-3, 1e-5     # src/main.py:325,367,365,194
SAMPLE = 1024


def sample_of(x):
    """Fixture-sized view of a tensor: a fixed-stride sample of at most SAMPLE values (the whole tensor if smaller)."""
    flat = np.asarray(x, dtype=np.float32).reshape(-1)
    step = max(1, flat.size // SAMPLE)
    return flat[::step][:SAMPLE].copy()


def run_train(ref_utils, RecurrentRGCN):
    """tests/golden/train_regcn.npz: TRAIN_STEPS optimisation steps of the UNMODIFIED reference (get_loss in train()
    mode with every dropout probability 0 -- torch's dropout generator cannot be reproduced by another implementation --
    backward, clip_grad_norm_, Adam) on the same history / triples.  Stored per case and step: the three losses, the
    clipped-from gradient norm, and per parameter the gradient norm + a strided sample of the gradient and of the
    updated value; after the last step the BatchNorm running statistics."""
    out = {}
    for name, cfg in TRAIN_CASES.items():
        case = synth.make_case(cfg["shape"], cfg["seed"])
        n, r = case["num_ents"], case["num_rels"]
        m = RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, H_DIM, "sub", 3, num_bases=N_BASES, num_basis=-1,
                          num_hidden_layers=N_LAYERS, dropout=0.0, self_loop=True,
                          skip_connect=cfg.get("skip_connect", False),
                          layer_norm=cfg["layer_norm"], input_dropout=0.0, hidden_dropout=0.0, feat_dropout=0.0,
                          entity_prediction=True, relation_prediction=True, use_cuda=False, gpu="cpu")
        m.load_state_dict(synth.fill_state_dict(m.state_dict(), cfg["seed"]))
        m.train()
        m.gpu = "cpu"
        # get_loss creates its accumulators as leaves and adds in place (src/rrgcn.py:205-207,219): that only works on
        # the use_cuda path, where .cuda() returns a (non-leaf) copy.  Mimic exactly that: .cuda() of a leaf that
        # requires grad is a differentiable copy, of anything else the identity (CPU stand-in for the device copy).
        torch.Tensor.cuda = lambda self, *a, **k: (self.clone() if (self.requires_grad and self.is_leaf) else self)
        opt = torch.optim.Adam(m.parameters(), lr=LR, weight_decay=WEIGHT_DECAY)
        glist = [ref_utils.build_sub_graph(n, r, snap, False, "cpu") for snap in case["history"]]
        triples = torch.from_numpy(case["test"])
        for step in range(TRAIN_STEPS):
            le, lr_, ls = m.get_loss(glist, triples.clone(), None, True)
            loss = TASK_WEIGHT * le + (1 - TASK_WEIGHT) * lr_ + ls
            loss.backward()
            tn = torch.nn.utils.clip_grad_norm_(m.parameters(), GRAD_NORM)
            grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}
            # clip_grad_norm_ scaled .grad in place: undo to store the raw gradient
            coef = min(1.0, GRAD_NORM / (float(tn) + 1e-6))
            opt.step()
            opt.zero_grad()
            out[f"{name}.s{step}.losses"] = np.array([float(le), float(lr_), float(ls)], dtype=np.float64)
            out[f"{name}.s{step}.grad_norm"] = np.array(float(tn), dtype=np.float64)
            for k, p in m.named_parameters():
                if grads[k] is None:
                    continue
                g = grads[k].numpy() / coef
                out[f"{name}.s{step}.gn.{k}"] = np.array(np.linalg.norm(g.astype(np.float64)))
                out[f"{name}.s{step}.g.{k}"] = sample_of(g)
                out[f"{name}.s{step}.p.{k}"] = sample_of(p.detach().numpy())
            print(name, step, [float(le), float(lr_)], "grad norm", float(tn))
        for k, v in m.state_dict().items():
            if "running_" in k and (".bn0." in k or ".bn1." in k or ".bn2." in k):
                out[f"{name}.bn.{k}"] = v.numpy().astype(np.float32)
        out[f"{name}.no_grad"] = np.array(json.dumps(sorted(k for k, p in m.named_parameters() if k not in
                                                              [kk for kk, g in grads.items() if g is not None])))
    path = os.path.join(GOLDEN, "train_regcn.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) / 1e6, "MB")


# ---------------------------------------------------------------------------------------------------------------
# static-graph constraint golden (src/rrgcn.py:101-106,146-152,225-247): predict + eval losses + one training step
# ---------------------------------------------------------------------------------------------------------------
STATIC_CASES = {
    "static_tiny_s0": dict(shape="tiny", seed=0, layer_norm=True, discount=1, angle=10, weight=0.5),
    "static_tiny_s1_noln": dict(shape="tiny", seed=1, layer_norm=False, discount=0, angle=10, weight=1.0),
    "static_small_s2": dict(shape="small", seed=2, layer_norm=True, discount=1, angle=10, weight=0.5),
}


def build_static_reference_model(cfg, n, r, n_srel, n_words, RecurrentRGCN, dropout=0.0):
    m = RecurrentRGCN("convtranse", "uvrgcn", n, r, n_srel, n_words, H_DIM, "sub", 3, num_bases=N_BASES, num_basis=-1,
                      num_hidden_layers=N_LAYERS, dropout=dropout, self_loop=True, skip_connect=False,
                      layer_norm=cfg["layer_norm"], input_dropout=dropout, hidden_dropout=dropout, feat_dropout=dropout,
                      weight=cfg["weight"], discount=cfg["discount"], angle=cfg["angle"], use_static=True,
                      entity_prediction=True, relation_prediction=True, use_cuda=False, gpu="cpu")
    m.load_state_dict(synth.fill_state_dict(m.state_dict(), cfg["seed"]))
    return m


def run_static(ref_utils, RecurrentRGCN):
    """tests/golden/train_static_regcn.npz: the UNMODIFIED reference with --add-static-graph semantics."""
    out = {}
    for name, cfg in STATIC_CASES.items():
        case = synth.make_case(cfg["shape"], cfg["seed"])
        n, r = case["num_ents"], case["num_rels"]
        st, n_srel, n_words = synth.make_static(n, cfg["seed"])
        m = build_static_reference_model(cfg, n, r, n_srel, n_words, RecurrentRGCN)
        m.gpu = "cpu"
        sg = ref_utils.build_sub_graph(n + n_words, n_srel, st, False, "cpu")
        glist = [ref_utils.build_sub_graph(n, r, snap, False, "cpu") for snap in case["history"]]
        triples = torch.from_numpy(case["test"])
        m.eval()
        with torch.no_grad():
            all_t, score, score_rel = m.predict(glist, r, sg, triples, False)
            hist, static_emb, h0, _, _ = m.forward(glist, sg, False)
            losses = m.get_loss(glist, triples.clone(), sg, False)
        out[f"{name}.score"] = score.numpy()
        out[f"{name}.score_rel"] = score_rel.numpy()
        out[f"{name}.static_emb"] = static_emb.numpy()
        out[f"{name}.hist_last"] = hist[-1].numpy()
        out[f"{name}.eval_losses"] = np.array([float(x.reshape(-1)[0]) for x in losses], dtype=np.float64)
        # one optimisation step (dropout 0), as in run_train
        m.train()
        torch.Tensor.cuda = lambda self, *a, **k: (self.clone() if (self.requires_grad and self.is_leaf) else self)
        opt = torch.optim.Adam(m.parameters(), lr=LR, weight_decay=WEIGHT_DECAY)
        le, lr_, ls = m.get_loss(glist, triples.clone(), sg, True)
        loss = TASK_WEIGHT * le + (1 - TASK_WEIGHT) * lr_ + ls
        loss.backward()
        tn = torch.nn.utils.clip_grad_norm_(m.parameters(), GRAD_NORM)
        coef = min(1.0, GRAD_NORM / (float(tn) + 1e-6))
        grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}
        opt.step()
        out[f"{name}.s0.losses"] = np.array([float(le.detach()), float(lr_.detach()), float(ls.detach())], dtype=np.float64)
        out[f"{name}.s0.grad_norm"] = np.array(float(tn), dtype=np.float64)
        for k, p in m.named_parameters():
            if grads[k] is None:
                continue
            g = grads[k].numpy() / coef
            out[f"{name}.s0.gn.{k}"] = np.array(np.linalg.norm(g.astype(np.float64)))
            out[f"{name}.s0.g.{k}"] = sample_of(g)
            out[f"{name}.s0.p.{k}"] = sample_of(p.detach().numpy())
        print(name, "eval losses", out[f"{name}.eval_losses"], "train losses", out[f"{name}.s0.losses"], "grad norm", float(tn))
    path = os.path.join(GOLDEN, "train_static_regcn.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) / 1e6, "MB")


def run_construct_snap(ref_utils):
    """tests/golden/aux_construct_snap.npz: the reference's construct_snap / construct_snap_r (rgcn/utils.py:367-405)
    on seeded score matrices without ties."""
    rng = np.random.default_rng(99)
    out = {}
    for name, (B, N, R, K) in {"ent": (40, 300, 7, 10), "rel": (40, 14, 7, 5)}.items():
        score = rng.permutation(B * N).reshape(B, N).astype(np.float32) / 7.0      # all distinct
        trip = np.stack([rng.integers(0, 300, B), rng.integers(0, 2 * R, B), rng.integers(0, 300, B)], 1).astype(np.int64)
        fn = ref_utils.construct_snap if name == "ent" else ref_utils.construct_snap_r
        res = fn(torch.from_numpy(trip), 300, R, torch.from_numpy(score), K)
        out[f"{name}.score"], out[f"{name}.triples"], out[f"{name}.out"] = score, trip, np.asarray(res, dtype=np.int64)
        out[f"{name}.cfg"] = np.array([B, N, R, K])
    np.savez_compressed(os.path.join(GOLDEN, "aux_construct_snap.npz"), **out)
    print("-> aux_construct_snap.npz")


# ---------------------------------------------------------------------------------------------------------------
# hyperbolic training golden (hyperbolic_main.py:585-628, hyperbolic_model.py:941-1088), dropout 0
# ---------------------------------------------------------------------------------------------------------------
HYP_TRAIN_CASES = {
    "hyptrain_tiny_s0": dict(kind="hyp", shape="tiny", seed=0, encoder="hyperbolic_uvrgcn", decoder="hyperbolic_convtranse",
                             layer_norm=False, gamma=0.15),
    "hyptrain_tiny_s1_ln": dict(kind="hyp", shape="tiny", seed=1, encoder="hyperbolic_uvrgcn",
                                decoder="hyperbolic_convtranse", layer_norm=True, gamma=1.0),
    "hyptrain_small_s2_ln": dict(kind="hyp", shape="small", seed=2, encoder="hyperbolic_uvrgcn",
                                 decoder="hyperbolic_convtranse", layer_norm=True, gamma=0.15),
    "hyptrain_murp_tiny_s3": dict(kind="hyp", shape="tiny", seed=3, encoder="hyperbolic_uvrgcn", decoder="murp",
                                  layer_norm=False, gamma=0.15),
    "hyptrain_murp_small_s4_bias": dict(kind="hyp", shape="small", seed=4, encoder="hyperbolic_uvrgcn", decoder="murp",
                                        layer_norm=True, gamma=0.15, entity_bias=True),
    "hyptrain_roth_tiny_s5": dict(kind="hyp", shape="tiny", seed=5, encoder="hyperbolic_uvrgcn", decoder="roth",
                                  layer_norm=False, gamma=0.15),
    "hyptrain_roth_small_s6_bias": dict(kind="hyp", shape="small", seed=6, encoder="hyperbolic_uvrgcn", decoder="roth",
                                        layer_norm=True, gamma=0.15, entity_bias=True),
    "hyptrain_atth_tiny_s7": dict(kind="hyp", shape="tiny", seed=7, encoder="hyperbolic_uvrgcn", decoder="atth",
                                  layer_norm=False, gamma=0.15),
    "hyptrain_atth_small_s8_bias": dict(kind="hyp", shape="small", seed=8, encoder="hyperbolic_uvrgcn", decoder="atth",
                                        layer_norm=True, gamma=0.15, entity_bias=True),
    "hyptrain_roth_small_s11_curv": dict(kind="hyp", shape="small", seed=11, encoder="hyperbolic_uvrgcn", decoder="roth",
                                         layer_norm=False, gamma=0.15, entity_bias=True, rel_curvature=True),
    "hyptrain_murp_tiny_s12_curv": dict(kind="hyp", shape="tiny", seed=12, encoder="hyperbolic_uvrgcn", decoder="murp",
                                        layer_norm=True, gamma=0.15, rel_curvature=True),
    "hyptrain_static_tiny_s13": dict(kind="hyp", shape="tiny", seed=13, encoder="hyperbolic_uvrgcn",
                                     decoder="hyperbolic_convtranse", layer_norm=True, gamma=0.15,
                                     static=dict(discount=1, angle=10, weight=0.5)),
    "hyptrain_static_small_s14_roth": dict(kind="hyp", shape="small", seed=14, encoder="hyperbolic_uvrgcn", decoder="roth",
                                           layer_norm=False, gamma=0.15, static=dict(discount=0, angle=10, weight=1.0)),
    "hyptrain_skip_tiny_s15": dict(kind="hyp", shape="tiny", seed=15, encoder="hyperbolic_uvrgcn",
                                   decoder="hyperbolic_convtranse", layer_norm=True, gamma=0.15, skip_connect=True),
    "hyptrain_lgcn_skip_small_s16": dict(kind="hyp", shape="small_l", seed=16, encoder="lgcn", decoder="roth",
                                         layer_norm=True, gamma=0.15, skip_connect=True),
    "hyptrain_lgcn_roth_small_s9": dict(kind="hyp", shape="small_l", seed=9, encoder="lgcn", decoder="roth",
                                        layer_norm=False, gamma=0.15),
    "hyptrain_lgcn_convtranse_small_s10_ln": dict(kind="hyp", shape="small_l", seed=10, encoder="lgcn",
                                                  decoder="hyperbolic_convtranse", layer_norm=True, gamma=0.15),
    # num_bases clamped to 2R = 10 (hyperbolic_layers.py:559-561): relation blocks of 20x20
    "hyptrain_lgcn_roth_tiny_l_s17": dict(kind="hyp", shape="tiny_l", seed=17, encoder="lgcn", decoder="roth",
                                          layer_norm=False, gamma=0.15),
    "hyptrain_lgcn_skip_tiny_l_s18_ln": dict(kind="hyp", shape="tiny_l", seed=18, encoder="lgcn",
                                             decoder="hyperbolic_convtranse", layer_norm=True, gamma=0.15,
                                             skip_connect=True),
}


def run_hyp_train(ref_utils, HyperbolicRecurrentRGCN, only=()):
    """tests/golden/train_hyp.npz: one optimisation step of the UNMODIFIED hyperbolic reference (get_loss in train() mode
    with every dropout 0, loss = 0.7 l_e + 0.3 l_r + l_static + l_radius, clip_grad_norm_(1.0), Adam).  With case names
    after --hyp-train only those are (re)generated and merged into the existing file."""
    out = {}
    if only:
        with np.load(os.path.join(GOLDEN, "train_hyp.npz")) as z:
            out = {k: z[k] for k in z.files if k.split(".")[0] not in only}
    for name, cfg in HYP_TRAIN_CASES.items():
        if only and name not in only:
            continue
        case = synth.make_case(cfg["shape"], cfg["seed"])
        n, r = case["num_ents"], case["num_rels"]
        st_cfg, sg, n_srel, n_words = cfg.get("static"), None, 0, 0
        if st_cfg:
            st, n_srel, n_words = synth.make_static(n, cfg["seed"])
            sg = ref_utils.build_sub_graph(n + n_words, n_srel, st, False, "cpu")
        m = HyperbolicRecurrentRGCN(cfg["decoder"], cfg["encoder"], n, r, n_srel, n_words, H_DIM, "sub", 3,
                                    num_bases=N_BASES,
                                    num_hidden_layers=N_LAYERS, dropout=0.0, c=CURV, self_loop=True,
                                    skip_connect=cfg.get("skip_connect", False),
                                    layer_norm=cfg["layer_norm"], input_dropout=0.0, hidden_dropout=0.0, feat_dropout=0.0,
                                    entity_prediction=True, relation_prediction=True, use_cuda=False, gpu="cpu",
                                    use_static=bool(st_cfg), **(st_cfg or {}),
                                    radius_msg_gamma=cfg["gamma"], hyp_init_scale=1e-3,
                                    use_entity_euclidean_bias=cfg.get("entity_bias", False),
                                    use_relation_specific_curvature=cfg.get("rel_curvature", False))
        m.load_state_dict(synth.fill_state_dict(m.state_dict(), cfg["seed"]))
        glist = [ref_utils.build_sub_graph(n, r, snap, False, "cpu") for snap in case["history"]]
        if st_cfg or cfg.get("skip_connect"):
            m.eval()
            with torch.no_grad():
                _, score, score_rel = m.predict(glist, r, sg, torch.from_numpy(case["test"]), False)
                hist, static_emb, _, _, _ = m.forward(glist, sg, False)
                ev = m.get_loss(glist, torch.from_numpy(case["test"]).clone(), sg, False)
            out[f"{name}.score"], out[f"{name}.score_rel"] = score.numpy(), score_rel.numpy()
            out[f"{name}.hist_last"] = hist[-1].numpy()
            if static_emb is not None:
                out[f"{name}.static_emb"] = static_emb.numpy()
            out[f"{name}.eval_losses"] = np.array([float(x.reshape(-1)[0]) for x in ev], dtype=np.float64)
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=LR, weight_decay=WEIGHT_DECAY)
        le, lr_, ls, lrad = m.get_loss(glist, torch.from_numpy(case["test"]), sg, False)
        loss = TASK_WEIGHT * le + (1 - TASK_WEIGHT) * lr_ + ls + lrad
        loss.backward()
        tn = torch.nn.utils.clip_grad_norm_(m.parameters(), GRAD_NORM)
        coef = min(1.0, GRAD_NORM / (float(tn) + 1e-6))
        grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}
        opt.step()
        out[f"{name}.s0.losses"] = np.array([float(x.detach().reshape(-1)[0]) for x in (le, lr_, ls, lrad)], dtype=np.float64)
        out[f"{name}.s0.grad_norm"] = np.array(float(tn), dtype=np.float64)
        for k, p in m.named_parameters():
            if grads[k] is None:
                continue
            g = grads[k].numpy() / coef
            out[f"{name}.s0.gn.{k}"] = np.array(np.linalg.norm(g.astype(np.float64)))
            out[f"{name}.s0.g.{k}"] = sample_of(g)
            out[f"{name}.s0.p.{k}"] = sample_of(p.detach().numpy())
        print(name, out[f"{name}.s0.losses"], "grad norm", float(tn), "params with grad", sum(g is not None for g in grads.values()))
    path = os.path.join(GOLDEN, "train_hyp.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) / 1e6, "MB")


def run_layers(ref_utils):
    """HyperbolicRGCNLayer (hyperbolic_src/hyperbolic_layers.py:21-161) on its own: the reference layer on seeded
    points of the ball, every structural variant (block sizes, self loop, skip gate, activation)."""
    import torch.nn.functional as F
    from hyperbolic_src.hyperbolic_layers import HyperbolicRGCNLayer
    out = {}
    variants = [dict(shape="tiny", nb=4, self_loop=True, skip=False, act=True, gamma=0.15),
                dict(shape="tiny_l", nb=10, self_loop=True, skip=True, act=True, gamma=1.0),
                dict(shape="small", nb=20, self_loop=False, skip=False, act=False, gamma=0.0),
                dict(shape="small_l", nb=100, self_loop=True, skip=True, act=True, gamma=0.15)]
    for k, v in enumerate(variants):
        case = synth.make_case(v["shape"], 40 + k)
        n, r = case["num_ents"], case["num_rels"]
        g = ref_utils.build_sub_graph(n, r, case["history"][0], False, "cpu")
        layer = HyperbolicRGCNLayer(H_DIM, H_DIM, 2 * r, v["nb"], c=CURV, activation=F.rrelu if v["act"] else None,
                                    self_loop=v["self_loop"], dropout=0.0, skip_connect=v["skip"],
                                    radius_msg_gamma=v["gamma"])
        layer.load_state_dict(synth.fill_state_dict(layer.state_dict(), 60 + k))
        layer.eval()
        rng = np.random.default_rng(900 + k)
        h = rng.standard_normal((n, H_DIM)).astype(np.float32)
        h = h / np.linalg.norm(h, axis=1, keepdims=True) * rng.uniform(0.2, 6.0, size=(n, 1)).astype(np.float32)
        prev = rng.standard_normal((n, H_DIM)).astype(np.float32)
        prev = prev / np.linalg.norm(prev, axis=1, keepdims=True) * rng.uniform(0.2, 6.0, size=(n, 1)).astype(np.float32)
        with torch.no_grad():
            y = layer(g, torch.from_numpy(h), None, torch.from_numpy(prev) if v["skip"] else None)
        out[f"v{k}_config"] = np.array(json.dumps(v))
        out[f"v{k}_out"] = y.numpy()
    path = os.path.join(GOLDEN, "aux_layer_hyp_rgcn.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) / 1e6, "MB")


def main(argv):
    ref_utils, RecurrentRGCN, HyperbolicRecurrentRGCN = _import_reference()
    if len(argv) > 1 and argv[1] == "--layers":
        return run_layers(ref_utils)
    torch.set_num_threads(os.cpu_count() or 1)
    if len(argv) > 1 and argv[1] == "--losses":
        return run_losses(ref_utils, RecurrentRGCN, HyperbolicRecurrentRGCN)
    if len(argv) > 1 and argv[1] == "--hyp-train":
        return run_hyp_train(ref_utils, HyperbolicRecurrentRGCN, tuple(argv[2:]))
    if len(argv) > 1 and argv[1] == "--construct-snap":
        return run_construct_snap(ref_utils)
    if len(argv) > 1 and argv[1] == "--static":
        return run_static(ref_utils, RecurrentRGCN)
    if len(argv) > 1 and argv[1] == "--train":
        return run_train(ref_utils, RecurrentRGCN)
    names = argv[1:] or list(CASES)
    for name in names:
        run_case(name, CASES[name], ref_utils, RecurrentRGCN, HyperbolicRecurrentRGCN)


if __name__ == "__main__":
    main(sys.argv)
