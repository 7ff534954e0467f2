"""CPU: the restatement oracle (oracle/restate.py) against the golden fixtures produced by the UNMODIFIED
reference (oracle/gen_golden.py).  This is what pins the oracle (the reference ships no tests of its own)."""
import numpy as np
import pytest
import torch

from oracle import restate, synth
from tests.helpers import CURV, N_BASES, build_model, close, golden_names, load_golden


def _graphs(case):
    return [restate.build_edges(s, case["num_ents"], case["num_rels"]) for s in case["history"]]


@pytest.mark.parametrize("name", golden_names())
def test_edge_index_matches_reference(name):
    cfg, z = load_golden(name)
    case = synth.make_case(cfg["shape"], cfg["seed"])
    R = case["num_rels"]
    for i, g in enumerate(_graphs(case)):
        assert np.array_equal(g["src"], z[f"g{i}_src"])
        assert np.array_equal(g["dst"], z[f"g{i}_dst"])
        assert np.array_equal(g["etype"], z[f"g{i}_type"])
        assert np.array_equal(g["norm"], z[f"g{i}_norm"])          # fp32 reciprocal of an int: exact
        rel_rowptr, rel_ents = restate.r2e(case["history"][i], R)
        uniq_r, r_len, r_to_e = z[f"g{i}_uniq_r"], z[f"g{i}_r_len"], z[f"g{i}_r_to_e"]
        present = np.nonzero(np.diff(rel_rowptr))[0]
        assert np.array_equal(uniq_r, np.concatenate((present, present + R)))
        for (b, e), r in zip(r_len, uniq_r):
            mine = rel_ents[rel_rowptr[r % R]:rel_rowptr[r % R + 1]]
            assert np.array_equal(np.sort(r_to_e[b:e]), mine)        # same *set* (the reference's order is a python set's)


@pytest.mark.parametrize("name", golden_names())
def test_restatement_matches_reference(name):
    cfg, z = load_golden(name)
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    _, sd = build_model(cfg, n, r)
    graphs = _graphs(case)
    torch.set_num_threads(8)
    with torch.no_grad():
        if cfg["kind"] == "regcn":
            all_t, score, score_rel, hist, h0 = restate.regcn_predict(sd, graphs, r, case["test"],
                                                                      layer_norm=cfg["layer_norm"])
        else:
            all_t, score, score_rel, hist, h0 = restate.hyp_predict(
                sd, graphs, r, case["test"], c=CURV, decoder=cfg["decoder"], layer_norm=cfg["layer_norm"],
                encoder=cfg["encoder"], gamma=cfg["gamma"], num_bases=min(N_BASES, 2 * r))
    assert np.array_equal(all_t, z["all_triples"])
    ok, worst = close(h0.numpy(), z["h0"])
    assert ok, f"h0 worst ratio {worst}"
    if "hist" in z:
        for i, h in enumerate(hist):
            ok, worst = close(h.numpy(), z["hist"][i])
            assert ok, f"hist[{i}] worst ratio {worst}"
        ok, worst = close(score.numpy(), z["score"], rtol=1e-4)
        assert ok, f"score worst ratio {worst}"
        ok, worst = close(score_rel.numpy(), z["score_rel"], rtol=1e-4)
        assert ok, f"score_rel worst ratio {worst}"
    else:
        rows, qrows = z["sub_rows"], z["sub_qrows"]
        ok, worst = close(hist[-1][rows].numpy(), z["hist_last_rows"])
        assert ok, f"hist_last worst ratio {worst}"
        ok, worst = close(hist[0][rows].numpy(), z["hist_first_rows"])
        assert ok, f"hist_first worst ratio {worst}"
        ok, worst = close(score[qrows][:, rows].numpy(), z["score_block"], rtol=1e-4)
        assert ok, f"score block worst ratio {worst}"
        ok, worst = close(score_rel[qrows].numpy(), z["score_rel_qrows"], rtol=1e-4)
        assert ok, f"score_rel worst ratio {worst}"
        if "score_full_rows" in z.files:                     # complete score rows of 16 queries (every candidate)
            ok, worst = close(score[z["full_qrows"]].numpy(), z["score_full_rows"], rtol=1e-4)
            assert ok, f"full score rows worst ratio {worst}"
    # ranks on the oracle's own scores: identical to the reference's except where fp32 noise flips a near-tie
    all_ans = synth.answers_of(case["test"], r, False)
    all_ans_r = synth.answers_of(case["test"], r, True)
    _, _, rank, frank = restate.total_rank(all_t, score.numpy(), all_ans, 0)
    _, _, rank_r, frank_r = restate.total_rank(all_t, score_rel.numpy(), all_ans_r, 1)
    for mine, ref, what in ((rank, z["rank"], "rank"), (frank, z["filter_rank"], "filter_rank"),
                            (rank_r, z["rank_rel"], "rank_rel"), (frank_r, z["filter_rank_rel"], "filter_rank_rel")):
        flips = np.mean(mine != ref)
        assert flips <= 0.02, f"{what}: {flips:.3%} of ranks differ from the reference"


@pytest.mark.parametrize("name", ["regcn_tiny_s0", "regcn_small_s2", "hyp_lgcn_roth_tiny_s0"])
def test_rank_restatement_exact_on_reference_scores(name):
    """Ranks recomputed from the *reference's own* score matrices must be bit-identical to the reference's ranks."""
    cfg, z = load_golden(name)
    case = synth.make_case(cfg["shape"], cfg["seed"])
    r = case["num_rels"]
    all_ans = synth.answers_of(case["test"], r, False)
    all_ans_r = synth.answers_of(case["test"], r, True)
    fm, m, rank, frank = restate.total_rank(z["all_triples"], z["score"], all_ans, 0)
    fmr, mr, rank_r, frank_r = restate.total_rank(z["all_triples"], z["score_rel"], all_ans_r, 1)
    assert np.array_equal(rank, z["rank"]) and np.array_equal(frank, z["filter_rank"])
    assert np.array_equal(rank_r, z["rank_rel"]) and np.array_equal(frank_r, z["filter_rank_rel"])
    np.testing.assert_allclose([fm, m, fmr, mr], z["mrr"], rtol=1e-6)


@pytest.mark.parametrize("name", [n for n in golden_names() if "c3" not in n and "c1" not in n])
def test_loss_restatement_matches_reference(name):
    """restate.regcn_loss / hyp_loss against the reference's get_loss() outputs (tests/golden/losses.json)."""
    import json
    import os
    from tests.helpers import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "losses.json")))[name]
    cfg, _ = load_golden(name)
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    _, sd = build_model(cfg, n, r)
    graphs = _graphs(case)
    with torch.no_grad():
        if cfg["kind"] == "regcn":
            mine = restate.regcn_loss(sd, graphs, r, case["test"], layer_norm=cfg["layer_norm"])
        else:
            mine = restate.hyp_loss(sd, graphs, r, case["test"], c=CURV, decoder=cfg["decoder"],
                                    layer_norm=cfg["layer_norm"], encoder=cfg["encoder"], gamma=cfg["gamma"],
                                    num_bases=min(N_BASES, 2 * r))
    np.testing.assert_allclose(mine, ref, rtol=1e-4, atol=1e-6)


# ----------------------------------------------------------------------------------------- training step (8f-1)
def _train_golden():
    import os
    from tests.helpers import GOLDEN
    return np.load(os.path.join(GOLDEN, "train_regcn.npz"))


from tests.helpers import compare_train_step  # noqa: E402


TRAIN_CASES = {"regcn_tiny_s0": ("tiny", 0, True), "regcn_tiny_s1_noln": ("tiny", 1, False),
               "regcn_small_s2": ("small", 2, True), "regcn_tiny_s3_skip": ("tiny", 3, True)}


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_oracle_train_step_matches_reference(name):
    """restate.regcn_train_steps (autograd over the restated forward + clip + Adam) against two optimisation steps of
    the UNMODIFIED reference (tests/golden/train_regcn.npz, oracle/gen_golden.py --train): losses, gradient norm,
    per-parameter gradients and updated values."""
    import regcn_b200 as R
    shape, seed, ln = TRAIN_CASES[name]
    z = _train_golden()
    case = synth.make_case(shape, seed)
    n, r = case["num_ents"], case["num_rels"]
    m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, 200, "sub", 3, num_bases=100, num_basis=-1,
                        num_hidden_layers=2, dropout=0.0, self_loop=True, skip_connect=name.endswith("_skip"),
                        layer_norm=ln, input_dropout=0.0, hidden_dropout=0.0, feat_dropout=0.0, entity_prediction=True,
                        relation_prediction=True, use_cuda=True, gpu=0)
    sd = synth.fill_state_dict(m.state_dict(), seed)
    graphs = [restate.build_edges(s, n, r) for s in case["history"]]
    log, bufs = restate.regcn_train_steps(sd, graphs, r, case["test"], layer_norm=ln, steps=2)
    for s, rec in enumerate(log):
        compare_train_step(z, name, s, rec["losses"], rec["grad_norm"], {k: v.numpy() for k, v in rec["grads"].items()},
                           {k: v.numpy() for k, v in rec["params"].items()})
    for k, v in bufs.items():
        if f"{name}.bn.{k}" not in z.files:
            continue                                                     # bn3 / bn_init are never used (src/decoder.py:73-76)
        np.testing.assert_allclose(v.numpy(), z[f"{name}.bn.{k}"], rtol=2e-3, atol=1e-5)   # after the 2nd step


@pytest.mark.parametrize("name", sorted(__import__("tests.helpers", fromlist=["STATIC_CASES"]).STATIC_CASES))
def test_oracle_static_graph_matches_reference(name):
    """Static-graph constraint (src/rrgcn.py:101-106,146-152,225-247): the restated static embedding, scores, eval
    losses and one optimisation step against the UNMODIFIED reference (tests/golden/train_static_regcn.npz)."""
    import os
    from tests.helpers import GOLDEN, STATIC_CASES, build_static_model
    z = np.load(os.path.join(GOLDEN, "train_static_regcn.npz"))
    cfg = STATIC_CASES[name]
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    st, n_srel, n_words = synth.make_static(n, cfg["seed"])
    _, sd = build_static_model(cfg, n, r, n_srel, n_words)
    graphs = [restate.build_edges(s, n, r) for s in case["history"]]
    sg = restate.build_edges(st, n + n_words, n_srel)
    ln = cfg["layer_norm"]
    P = {k: v for k, v in sd.items() if v.is_floating_point()}
    with torch.no_grad():
        s_emb = restate.regcn_static_emb(P, sg, n, 100, ln)
        all_t, score, score_rel, hist, h0 = restate.regcn_predict(sd, graphs, r, case["test"], layer_norm=ln, h_init=s_emb)
        l_static = restate.static_angle_loss(s_emb, hist, ln, cfg["angle"], cfg["discount"], cfg["weight"])
    for mine, key in ((s_emb, "static_emb"), (hist[-1], "hist_last"), (score, "score"), (score_rel, "score_rel")):
        ok, worst = close(mine.numpy(), z[f"{name}.{key}"], rtol=1e-4)
        assert ok, (key, worst)
    np.testing.assert_allclose(float(l_static), z[f"{name}.eval_losses"][2], rtol=1e-4)
    static = dict(graph=sg, num_ents=n, num_bases=100, angle=cfg["angle"], discount=cfg["discount"], weight=cfg["weight"])
    log, _ = restate.regcn_train_steps(sd, graphs, r, case["test"], layer_norm=ln, steps=1, static=static)
    rec = log[0]
    compare_train_step(z, name, 0, rec["losses"], rec["grad_norm"], {k: v.numpy() for k, v in rec["grads"].items()},
                       {k: v.numpy() for k, v in rec["params"].items()})


def test_oracle_construct_snap_matches_reference():
    """Multi-step feedback (rgcn/utils.py:367-405) against the reference's own outputs (aux_construct_snap.npz)."""
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "aux_construct_snap.npz"))
    for name, mode in (("ent", 0), ("rel", 1)):
        B, N, Rr, K = (int(v) for v in z[f"{name}.cfg"])
        got = restate.construct_snap(z[f"{name}.triples"], Rr, z[f"{name}.score"], K, mode)
        assert np.array_equal(got, z[f"{name}.out"])


@pytest.mark.parametrize("name", sorted(__import__("tests.helpers", fromlist=["HYP_TRAIN_CASES"]).HYP_TRAIN_CASES))
def test_oracle_hyperbolic_train_step_matches_reference(name):
    """restate.hyp_train_steps against one optimisation step of the UNMODIFIED hyperbolic reference
    (tests/golden/train_hyp.npz, oracle/gen_golden.py --hyp-train)."""
    import os
    from tests.helpers import GOLDEN, HYP_TRAIN_CASES, build_hyp_train_model
    z = np.load(os.path.join(GOLDEN, "train_hyp.npz"))
    cfg = HYP_TRAIN_CASES[name]
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    _, sd = build_hyp_train_model(cfg, n, r)
    graphs = [restate.build_edges(s, n, r) for s in case["history"]]
    static = None
    if cfg.get("static"):
        # --add-static-graph (hyperbolic_src/hyperbolic_model.py:762-771,1039-1064)
        st, n_srel, n_words = synth.make_static(n, cfg["seed"])
        static = dict(graph=restate.build_edges(st, n + n_words, n_srel), num_ents=n, num_bases=N_BASES, **cfg["static"])
        P = {k: v for k, v in sd.items() if v.is_floating_point()}
        with torch.no_grad():
            s_emb = restate.hyp_static_emb(P, static["graph"], n, N_BASES, cfg["layer_norm"])
            ls = restate.hyp_train_losses(P, graphs, r, case["test"], CURV, cfg["layer_norm"], cfg["gamma"], {},
                                          decoder=cfg["decoder"], encoder=cfg["encoder"],
                                          num_bases=min(N_BASES, 2 * r), static=static)
        np.testing.assert_allclose(s_emb.numpy(), z[f"{name}.static_emb"], rtol=1e-4, atol=1e-6)
    log = restate.hyp_train_steps(sd, graphs, r, case["test"], c=CURV, layer_norm=cfg["layer_norm"], gamma=cfg["gamma"],
                                  decoder=cfg["decoder"], encoder=cfg["encoder"], num_bases=min(N_BASES, 2 * r),
                                  static=static)
    rec = log[0]
    compare_train_step(z, name, 0, rec["losses"], rec["grad_norm"], {k: v.numpy() for k, v in rec["grads"].items()},
                       {k: v.numpy() for k, v in rec["params"].items()})


def test_rna_tf32_restatement_matches_bit_pattern_rule():
    """oracle.restate.rna_tf32 (frexp / floor in float64) == add-half-ulp-and-mask on the bit pattern (what the kernels
    compute) on normal fp32 values incl. exact ties."""
    import numpy as np
    from oracle import restate
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-40, 40, 100000))).astype(np.float32)
    ties = (np.arange(1, 4097, dtype=np.uint32) << 13 | 0x1000 | 0x3f800000).view(np.float32)
    x = np.concatenate([x, ties, -ties])
    bits = ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
    assert np.array_equal(bits, restate.rna_tf32(x))


def test_hyp_rgcn_layer_restatement_matches_reference():
    """restate.hyp_rgcn_layer == the reference's HyperbolicRGCNLayer (hyperbolic_layers.py:21-161) on every structural
    variant of tests/golden/aux_layer_hyp_rgcn.npz (block sizes, self loop, skip gate, activation)."""
    import regcn_b200.hyperbolic_layers as HL
    import torch.nn.functional as F
    from tests.helpers import hyp_rgcn_layer_cases
    n_cases = 0
    for k, v, case, h, prev, want in hyp_rgcn_layer_cases():
        n, r = case["num_ents"], case["num_rels"]
        layer = HL.HyperbolicRGCNLayer(200, 200, 2 * r, v["nb"], c=CURV, activation=F.rrelu if v["act"] else None,
                                       self_loop=v["self_loop"], skip_connect=v["skip"], radius_msg_gamma=v["gamma"])
        sd = synth.fill_state_dict(layer.state_dict(), 60 + k)
        g = restate.build_edges(case["history"][0], n, r)
        got = restate.hyp_rgcn_layer(torch.from_numpy(h), g, sd["weight"], layer.num_bases, CURV, v["gamma"],
                                     w_loop=sd.get("loop_weight"),
                                     skip=(sd["skip_weight"], sd["skip_bias"]) if v["skip"] else None,
                                     prev_h=torch.from_numpy(prev) if v["skip"] else None, act=v["act"])
        ok, worst = close(got.numpy(), want, rtol=1e-5)
        assert ok, (k, worst)
        n_cases += 1
    assert n_cases == 4
