"""GPU parity of the training step: every autograd node of regcn_b200/train.py against torch
autograd over the CPU oracle (oracle/restate.py), the whole optimisation step against two steps of the UNMODIFIED
reference (tests/golden/train_regcn.npz), and the fused clipped Adam against torch.optim.Adam.

Tolerances: gradients |g - g_ref| <= 2e-4 * max(max|g_ref|, 1e-3 * |all gradients|) (tests/helpers.grad_close), forward
values 1e-4 * max(1, |ref|)."""
import numpy as np
import pytest
import torch

import regcn_b200 as R
from oracle import restate, synth
from regcn_b200 import optim, train
from tests.helpers import close, compare_train_step, grad_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _leaf(a, dev=DEV):
    return torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev).requires_grad_(True)


def _cmp_grads(mine, ref, names, rtol=2e-4):
    tn = float(np.sqrt(sum(float((r.double() ** 2).sum()) for r in ref)))
    for m, r, n in zip(mine, ref, names):
        ok, worst = grad_close(m.detach().cpu().numpy(), r.detach().numpy(), tn, rtol)
        assert ok, (n, worst)


def _graph(shape="small", seed=3):
    case = synth.make_case(shape, seed)
    n, r = case["num_ents"], case["num_rels"]
    g = R.build_sub_graph(n, r, case["history"][0], True, 0)
    og = restate.build_edges(case["history"][0], n, r)
    return case, n, r, g, og


@pytest.mark.parametrize("w_kn", [True, False])
def test_linear_forward_backward(w_kn):
    R._lib.require_device()
    rng = np.random.default_rng(0)
    M, K, N = 37, 200, 400
    x = rng.standard_normal((M, K))
    W = rng.standard_normal((K, N) if w_kn else (N, K)) * 0.1
    b = rng.standard_normal(N)
    go = rng.standard_normal((M, N))
    xd, Wd, bd = _leaf(x), _leaf(W), _leaf(b)
    y = train.linear(xd, Wd, bd, w_kn)
    y.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    xc, Wc, bc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, W, b))
    yc = (xc @ Wc if w_kn else xc @ Wc.t()) + bc
    yc.backward(torch.as_tensor(go))
    ok, worst = close(y.detach().cpu().numpy(), yc.detach().numpy())
    assert ok, worst
    _cmp_grads([xd.grad, Wd.grad, bd.grad], [xc.grad, Wc.grad, bc.grad], ["dx", "dW", "db"])


def test_group_by_key_and_gather_sum():
    R._lib.require_device()
    rng = np.random.default_rng(1)
    n, nkeys, d = 1000, 37, 200
    keys = rng.integers(0, nkeys, size=n).astype(np.int32)
    keys[keys == 5] = 6                                           # an empty group
    vals = rng.integers(0, 500, size=n).astype(np.int32)
    rp, perm, vout = train._group(torch.from_numpy(keys).to(DEV), nkeys, torch.from_numpy(vals).to(DEV))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(perm.cpu().numpy()[:n], order)
    assert np.array_equal(vout.cpu().numpy()[:n], vals[order])
    assert np.array_equal(rp.cpu().numpy(), np.searchsorted(keys[order], np.arange(nkeys + 1)))
    X = rng.standard_normal((500, d)).astype(np.float32)
    w = rng.random(500).astype(np.float32)
    out = train._gather_sum(torch.from_numpy(X).to(DEV), d, torch.from_numpy(w).to(DEV), rp, vout, nkeys, d)
    ref = np.zeros((nkeys, d))
    np.add.at(ref, keys, X[vals].astype(np.float64) * w[vals][:, None])
    ok, worst = close(out.cpu().numpy(), ref)
    assert ok, worst


def test_edge_path_nodes_backward():
    """rel_mean_pool, union_aggregate, union_combine, time_gate, gru_gate, normalize, tanh against autograd on the oracle."""
    R._lib.require_device()
    case, n, r, g, og = _graph()
    rng = np.random.default_rng(2)
    d = 200
    h = rng.standard_normal((n, d))
    rel = rng.standard_normal((2 * r, d)) * 0.3
    # --- mean pool
    hd = _leaf(h)
    go = rng.standard_normal((2 * r, d))
    train.rel_mean_pool(hd, g).backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    hc = torch.tensor(h, dtype=torch.float64, requires_grad=True)
    rp, ents = restate.r2e(og["triples"], r)
    restate.rel_mean_pool(hc, rp, ents, r).backward(torch.as_tensor(go))
    _cmp_grads([hd.grad], [hc.grad], ["meanpool dh"])
    # --- aggregate
    hd, rd = _leaf(h), _leaf(rel)
    go = rng.standard_normal((n, d))
    agg = train.union_aggregate(hd, rd, g)
    agg.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    hc = torch.tensor(h, dtype=torch.float64, requires_grad=True)
    rc = torch.tensor(rel, dtype=torch.float64, requires_grad=True)
    msg = hc[torch.as_tensor(og["src"])] + rc[torch.as_tensor(og["etype"])]
    aggc = restate.scatter_sum(msg, og["dst"], n) * torch.as_tensor(og["norm"]).double().view(-1, 1)
    aggc.backward(torch.as_tensor(go))
    ok, worst = close(agg.detach().cpu().numpy(), aggc.detach().numpy())
    assert ok, worst
    _cmp_grads([hd.grad, rd.grad], [hc.grad, rc.grad], ["agg dh", "agg drel"])
    # --- combine (p = 0)
    P, L = rng.standard_normal((n, d)), rng.standard_normal((n, 2 * d))
    Pd, Ld = _leaf(P), _leaf(L)
    out = train.union_combine(Pd, Ld, g, 0.0)
    out.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    Pc, Lc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (P, L))
    has_in = torch.as_tensor(og["indeg"] > 0).view(-1, 1)
    outc = restate.rrelu(Pc + torch.where(has_in, Lc[:, :d], Lc[:, d:]))
    outc.backward(torch.as_tensor(go))
    ok, worst = close(out.detach().cpu().numpy(), outc.detach().numpy())
    assert ok, worst
    _cmp_grads([Pd.grad, Ld.grad], [Pc.grad, Lc.grad], ["combine dP", "combine dL"])
    # --- time gate, with and without the normalisation
    for norm in (True, False):
        G, b, cur = rng.standard_normal((n, d)), rng.standard_normal(d) * 0.1, rng.standard_normal((n, d))
        Gd, bd, cd, hd = _leaf(G), _leaf(b), _leaf(cur), _leaf(h)
        o = train.time_gate(Gd, bd, cd, hd, norm)
        o.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
        Gc, bc, cc, hc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (G, b, cur, h))
        tw = torch.sigmoid(Gc + bc)
        oc = tw * (restate.normalize_rows(cc) if norm else cc) + (1 - tw) * hc
        oc.backward(torch.as_tensor(go))
        ok, worst = close(o.detach().cpu().numpy(), oc.detach().numpy())
        assert ok, worst
        _cmp_grads([Gd.grad, bd.grad, cd.grad, hd.grad], [Gc.grad, bc.grad, cc.grad, hc.grad], ["dG", "db", "dcur", "dh"])
    # --- GRU gates
    for norm in (True, False):
        M = 2 * r
        gi, gh, hp = rng.standard_normal((M, 3 * d)), rng.standard_normal((M, 3 * d)), rng.standard_normal((M, d))
        go2 = rng.standard_normal((M, d))
        gid, ghd, hpd = _leaf(gi), _leaf(gh), _leaf(hp)
        o = train.gru_gate(gid, ghd, hpd, norm)
        o.backward(torch.as_tensor(go2, dtype=torch.float32, device=DEV))
        gic, ghc, hpc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (gi, gh, hp))
        rr = torch.sigmoid(gic[:, :d] + ghc[:, :d])
        zz = torch.sigmoid(gic[:, d:2 * d] + ghc[:, d:2 * d])
        nn_ = torch.tanh(gic[:, 2 * d:] + rr * ghc[:, 2 * d:])
        oc = (hpc - nn_) * zz + nn_
        oc = restate.normalize_rows(oc) if norm else oc
        oc.backward(torch.as_tensor(go2))
        ok, worst = close(o.detach().cpu().numpy(), oc.detach().numpy())
        assert ok, worst
        _cmp_grads([gid.grad, ghd.grad, hpd.grad], [gic.grad, ghc.grad, hpc.grad], ["dgi", "dgh", "dhprev"])
    # --- normalize, tanh
    for fn, ref_fn in ((train.normalize, restate.normalize_rows), (train.tanh, torch.tanh)):
        hd = _leaf(h)
        fn(hd).backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
        hc = torch.tensor(h, dtype=torch.float64, requires_grad=True)
        ref_fn(hc).backward(torch.as_tensor(go))
        _cmp_grads([hd.grad], [hc.grad], [fn.__name__ if hasattr(fn, "__name__") else "row"])


def _decoder_pair(n, r, seed, p=0.0):
    m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, 200, "sub", 3, num_bases=100, num_basis=-1,
                        num_hidden_layers=2, dropout=p, self_loop=True, skip_connect=False, layer_norm=True,
                        input_dropout=p, hidden_dropout=p, feat_dropout=p, entity_prediction=True,
                        relation_prediction=True, use_cuda=True, gpu=0)
    sd = synth.fill_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    return m.to(DEV).train(), sd


@pytest.mark.parametrize("which", ["decoder_ob", "rdecoder"])
def test_conv_tower_and_ce_backward(which):
    """Train-mode ConvTransE / ConvTransR tower (batch-statistics BatchNorm) + all-candidate cross entropy."""
    R._lib.require_device()
    case = synth.make_case("small", 5)
    n, r = case["num_ents"], case["num_rels"]
    m, sd = _decoder_pair(n, r, 5)
    mod = getattr(m, which)
    rng = np.random.default_rng(7)
    d = 200
    e_all = np.tanh(rng.standard_normal((n, d)))
    rel = rng.standard_normal((2 * r, d)) * 0.3
    all_t = restate.add_inverse(case["test"], r)
    t_dev = torch.from_numpy(all_t).to(DEV)
    ed, rd = _leaf(e_all), _leaf(rel)
    if which == "decoder_ob":
        q = train.conv_tower(mod, ed, rd, t_dev, 0, 1)
        loss = train.score_ce(q, ed, t_dev, 2)
    else:
        q = train.conv_tower(mod, ed, ed, t_dev, 0, 2)
        loss = train.score_ce(q, rd, t_dev, 1)
    loss.backward()
    P = {k: v.clone().double() for k, v in sd.items() if v.is_floating_point()}
    names = [k for k in P if k.startswith(which + ".") and "running" not in k and ".bn3." not in k and ".bn_init." not in k
             and not k.endswith(".b")]
    for k in names:
        P[k].requires_grad_(True)
    ec = torch.tensor(e_all, dtype=torch.float64, requires_grad=True)
    rc = torch.tensor(rel, dtype=torch.float64, requires_grad=True)
    tt = torch.as_tensor(all_t)
    stats = {}
    if which == "decoder_ob":
        qc = restate.conv_tower_train(ec[tt[:, 0]], rc[tt[:, 1]], P, which + ".", stats)
        lc = restate.cross_entropy(qc @ ec.t(), all_t[:, 2])
    else:
        qc = restate.conv_tower_train(ec[tt[:, 0]], ec[tt[:, 2]], P, which + ".", stats)
        lc = restate.cross_entropy(qc @ rc.t(), all_t[:, 1])
    lc.backward()
    ok, worst = close(q.detach().cpu().numpy(), qc.detach().numpy())
    assert ok, worst
    assert abs(float(loss) - float(lc)) <= 1e-4 * max(1.0, abs(float(lc)))
    mine = [ed.grad, rd.grad] + [dict(m.named_parameters())[k].grad for k in names]
    ref = [ec.grad, rc.grad] + [P[k].grad for k in names]
    assert all(g is not None for g in mine)
    _cmp_grads(mine, ref, ["d e_all", "d rel"] + names)
    for k, v in stats.items():                       # running statistics updated like nn.BatchNorm1d
        np.testing.assert_allclose(m.state_dict()[k].cpu().numpy(), v.numpy(), rtol=1e-4, atol=1e-6)


def test_adam_matches_torch():
    R._lib.require_device()
    rng = np.random.default_rng(11)
    shapes = [(64, 200), (200,), (50, 2, 3), (7,)]
    ps = [torch.nn.Parameter(torch.tensor(rng.standard_normal(s), dtype=torch.float32, device=DEV)) for s in shapes]
    pc = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ps]
    unused = torch.nn.Parameter(torch.ones(5, device=DEV))
    opt = optim.Adam(ps + [unused], lr=1e-3, weight_decay=1e-5)
    ref = torch.optim.Adam(pc, lr=1e-3, weight_decay=1e-5)
    for step in range(3):
        gs = [rng.standard_normal(s) * (10.0 if step == 0 else 0.01) for s in shapes]      # clipped, then not clipped
        for p, c, g in zip(ps, pc, gs):
            gt = torch.tensor(g, dtype=torch.float32)
            if p.grad is None:
                p.grad = gt.to(DEV)
            else:
                p.grad.copy_(gt)
            c.grad = gt.clone()
        tn_ref = torch.nn.utils.clip_grad_norm_(pc, 1.0)
        tn = optim.clip_grad_norm_(opt, 1.0)
        opt.step()
        ref.step()
        np.testing.assert_allclose(float(opt.total_norm), float(tn_ref), rtol=1e-5)
        for p, c in zip(ps, pc):
            np.testing.assert_allclose(p.detach().cpu().numpy(), c.detach().numpy(), rtol=1e-5, atol=2e-7)
        opt.zero_grad()
        assert all(float(p.grad.abs().max()) == 0.0 for p in ps)
    assert torch.equal(unused.detach().cpu(), torch.ones(5))        # no gradient -> untouched, like torch.optim.Adam


@pytest.mark.parametrize("name,shape,seed,ln", [("regcn_tiny_s0", "tiny", 0, True), ("regcn_tiny_s1_noln", "tiny", 1, False),
                                                ("regcn_small_s2", "small", 2, True),
                                                ("regcn_tiny_s3_skip", "tiny", 3, True)])
def test_train_steps_match_reference(name, shape, seed, ln):
    """Two optimisation steps (get_loss in train() mode -> backward -> clip_grad_norm_(1.0) -> Adam) against the
    UNMODIFIED reference's (tests/golden/train_regcn.npz): losses, gradient norm, every gradient, every updated value,
    BatchNorm running statistics."""
    import os
    from tests.helpers import GOLDEN
    R._lib.require_device()
    z = np.load(os.path.join(GOLDEN, "train_regcn.npz"))
    case = synth.make_case(shape, seed)
    n, r = case["num_ents"], case["num_rels"]
    m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, 200, "sub", 3, num_bases=100, num_basis=-1,
                        num_hidden_layers=2, dropout=0.0, self_loop=True, skip_connect=name.endswith("_skip"),
                        layer_norm=ln, input_dropout=0.0, hidden_dropout=0.0, feat_dropout=0.0, entity_prediction=True,
                        relation_prediction=True, use_cuda=True, gpu=0)
    m.load_state_dict(synth.fill_state_dict(m.state_dict(), seed))
    m = m.to(DEV).train()
    opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    triples = torch.from_numpy(case["test"]).to(DEV)
    assert m._engine_ok()            # --skip-connect is dead for uvrgcn (prev_h=[], src/rrgcn.py:37-38): engine path too
    for step in range(2):
        le, lr_, ls = m.get_loss(glist, triples, None, True)
        loss = 0.7 * le + 0.3 * lr_ + ls
        loss.backward()
        named = {k: p for k, p in m.named_parameters() if p.grad is not None}
        grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in named.items()}
        optim.clip_grad_norm_(opt, 1.0)
        opt.step()
        params = {k: p.detach().cpu().numpy().copy() for k, p in named.items()}
        opt.zero_grad()
        # single nodes hold 2e-4 (tests above).  Through the whole chain (~45 GEMMs deep, un-normalised in the noln case)
        # cancellation amplifies the 22-mantissa-bit 3xTF32 operand split: measured against the fp64 oracle on B200 the
        # kernel path is within 1.2e-3 of each gradient's largest element (the fp32 reference: 2e-6 .. 1e-3), so the
        # whole-step gate is 1.5e-3 (+ the ReLU-kink floor of helpers.grad_close)
        compare_train_step(z, name, step, (float(le), float(lr_)), float(opt.total_norm), grads, params, rtol=1.5e-3)
    sd = m.state_dict()
    for k in z.files:
        if k.startswith(name + ".bn."):
            # after the SECOND step the two runs sit at slightly different points (helpers.compare_train_step)
            np.testing.assert_allclose(sd[k[len(name) + 4:]].cpu().numpy(), z[k], rtol=2e-3, atol=5e-4)
    # the trained parameters feed the evaluation engine (operand caches key on the version counter)
    m.eval()
    _, score, _ = m.predict(glist, r, None, triples, True)
    assert torch.isfinite(score).all()


def test_dropout_statistics_and_determinism():
    """Train-mode dropout: keep rate 1-p, survivors scaled by 1/(1-p), same seed -> same step, and a finite loss that
    decreases over a few steps on a fixed batch."""
    R._lib.require_device()
    x = torch.ones(1 << 20, device=DEV)
    R._lib.call("regcn_dropout", x.data_ptr(), x.numel(), 0.2, 1234)
    keep = float((x != 0).float().mean())
    assert abs(keep - 0.8) < 5e-3
    assert float(x.max()) == pytest.approx(1.25)
    case = synth.make_case("small", 9)
    n, r = case["num_ents"], case["num_rels"]
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    triples = torch.from_numpy(case["test"]).to(DEV)
    runs = []
    for _ in range(2):
        m, _sd = _decoder_pair(n, r, 9, p=0.2)
        opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
        train.manual_seed(77)
        losses = []
        for _step in range(6):
            le, lr_, ls = m.get_loss(glist, triples, None, True)
            (0.7 * le + 0.3 * lr_ + ls).backward()
            optim.clip_grad_norm_(opt, 1.0)
            opt.step()
            opt.zero_grad()
            losses.append(float(le))
        runs.append(losses)
    assert runs[0] == runs[1]                                  # bit-reproducible: fixed-order reductions, counter RNG
    assert np.isfinite(runs[0]).all() and runs[0][-1] < runs[0][0]


# ----------------------------------------------------------------------------------------- static-graph constraint (8f-3)
def test_block_layer_and_angle_loss_backward():
    """RGCNBlockLayer aggregate backward (dh, dW) and the angle loss gradient against autograd on the oracle."""
    R._lib.require_device()
    n, n_srel, n_words, d = 300, 3, 20, 200
    st, _, _ = synth.make_static(n, 4, n_srel, n_words)
    g = R.build_sub_graph(n + n_words, n_srel, st, True, 0)
    og = restate.build_edges(st, n + n_words, n_srel)
    rng = np.random.default_rng(12)
    h = rng.standard_normal((n + n_words, d))
    W = rng.standard_normal((2 * n_srel, 400)) * 0.3
    go = rng.standard_normal((n + n_words, d))
    hd, Wd = _leaf(h), _leaf(W)
    out = train._RReluDrop.apply(train._BlockAggregate.apply(hd, Wd, g, 100, d), 0.0)
    out.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    hc, Wc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (h, W))
    outc = restate.block_layer(hc, og, Wc, 100, d)
    outc.backward(torch.as_tensor(go))
    ok, worst = close(out.detach().cpu().numpy(), outc.detach().numpy())
    assert ok, worst
    _cmp_grads([hd.grad, Wd.grad], [hc.grad, Wc.grad], ["block dh", "block dW"])
    for ln, discount in ((True, 1), (False, 0)):
        s = rng.standard_normal((n, d))
        s = s / np.linalg.norm(s, axis=1, keepdims=True) if ln else s
        hist = [rng.standard_normal((n, d)) * (0.2 if t == 0 else 1.0) + (3.0 * s if t == 1 else 0.0) for t in range(3)]
        sd_, hd_ = _leaf(s), [_leaf(e) for e in hist]
        loss = train._StaticAngle.apply(sd_, ln, 10.0, discount, 0.5, *hd_)
        (loss * 1.7).backward()
        sc = torch.tensor(s, dtype=torch.float64, requires_grad=True)
        hc_ = [torch.tensor(e, dtype=torch.float64, requires_grad=True) for e in hist]
        lc = restate.static_angle_loss(sc, hc_, ln, 10.0, discount, 0.5)
        (lc * 1.7).backward()
        assert abs(float(loss.detach()) - float(lc.detach())) <= 1e-4 * max(1.0, abs(float(lc.detach())))
        _cmp_grads([sd_.grad] + [e.grad for e in hd_], [sc.grad] + [e.grad for e in hc_], ["dstatic", "de0", "de1", "de2"])


@pytest.mark.parametrize("name", ["static_tiny_s0", "static_tiny_s1_noln", "static_small_s2"])
def test_static_graph_model_matches_reference(name):
    """use_static=True end to end against the UNMODIFIED reference (tests/golden/train_static_regcn.npz): static
    embedding, evolved table and scores through the one-call engine, the three eval-mode losses, then one optimisation
    step (losses, every gradient incl. words_emb / statci_rgcn_layer.weight, updated values)."""
    import os
    from tests.helpers import GOLDEN, STATIC_CASES, build_static_model
    R._lib.require_device()
    z = np.load(os.path.join(GOLDEN, "train_static_regcn.npz"))
    cfg = STATIC_CASES[name]
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    st, n_srel, n_words = synth.make_static(n, cfg["seed"])
    m, _ = build_static_model(cfg, n, r, n_srel, n_words)
    m = m.to(DEV).eval()
    sg = R.build_sub_graph(n + n_words, n_srel, st, True, 0)
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    triples = torch.from_numpy(case["test"]).to(DEV)
    assert m._engine_ok()
    _, score, score_rel = m.predict(glist, r, sg, triples, True)
    hist, static_emb, _, _, _ = m.forward(glist, sg, True)
    for mine, key in ((static_emb, "static_emb"), (hist[-1], "hist_last"), (score, "score"), (score_rel, "score_rel")):
        ok, worst = close(mine.cpu().numpy(), z[f"{name}.{key}"], rtol=1e-4)
        assert ok, (key, worst)
    losses = m.get_loss(glist, triples, sg, True)
    np.testing.assert_allclose([float(x.reshape(-1)[0]) for x in losses], z[f"{name}.eval_losses"], rtol=1e-4)
    m.train()
    opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    le, lr_, ls = m.get_loss(glist, triples, sg, True)
    (0.7 * le + 0.3 * lr_ + ls).backward()
    named = {k: p for k, p in m.named_parameters() if p.grad is not None}
    grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in named.items()}
    optim.clip_grad_norm_(opt, 1.0)
    opt.step()
    params = {k: p.detach().cpu().numpy().copy() for k, p in named.items()}
    compare_train_step(z, name, 0, (float(le.detach()), float(lr_.detach()), float(ls.detach())), float(opt.total_norm),
                       grads, params, rtol=1.5e-3)


def test_fit_epoch_equals_hand_written_loop_and_learns():
    """fit_epoch (src/main.py:213-246 as a library call: snapshot cache, one sync per epoch) against the same steps
    written out by hand -- identical losses (the step is bit-reproducible) -- and the loss goes down over epochs."""
    R._lib.require_device()
    st = synth.make_stream("small", 11, n_test=5)
    n, r = st["num_ents"], st["num_rels"]
    train_list = st["history"] + st["tests"]
    L = 3

    def fresh():
        m, _ = _decoder_pair(n, r, 11, p=0.2)
        return m, optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)

    order = [3, 1, 0, 5, 2, 7, 4, 6]
    m1, o1 = fresh()
    train.manual_seed(5)
    res = R.fit_epoch(m1, o1, train_list, r, n, L, task_weight=0.7, grad_norm=1.0, order=order)
    m2, o2 = fresh()
    train.manual_seed(5)
    hand = []
    for t in order:
        if t == 0:
            continue
        glist = [R.build_sub_graph(n, r, s, True, 0) for s in train_list[max(0, t - L):t]]
        le, lr_, ls = m2.get_loss(glist, torch.from_numpy(train_list[t]).to(DEV), None, True)
        loss = 0.7 * le + 0.3 * lr_ + ls
        loss.backward()
        optim.clip_grad_norm_(o2, 1.0)
        o2.step()
        o2.zero_grad()
        hand.append(float(loss.detach()))
    assert res["steps"] == len(hand) == 7
    np.testing.assert_allclose(res["per_step"], hand, rtol=1e-6)
    first = res["loss"]
    for _ in range(4):
        res = R.fit_epoch(m1, o1, train_list, r, n, L, order=order)
    assert np.isfinite(res["loss"]) and res["loss"] < first
    # hyperbolic model with the reference's mini-batch accumulation (hyperbolic_main.py:585-598)
    from tests.helpers import build_hyp_train_model
    cfg = dict(kind="hyp", shape="small", seed=3, encoder="hyperbolic_uvrgcn", decoder="hyperbolic_convtranse",
               layer_norm=True, gamma=0.15)
    mh, _ = build_hyp_train_model(cfg, n, r, dropout=0.2)
    mh = mh.to(DEV)
    oh = optim.Adam(mh.parameters(), lr=1e-3, weight_decay=1e-5)
    a = R.fit_epoch(mh, oh, train_list, r, n, L, order=order, triple_batch_size=64)
    for _ in range(2):
        b = R.fit_epoch(mh, oh, train_list, r, n, L, order=order, triple_batch_size=64)
    assert a["steps"] == 7 and "loss_radius" in a and np.isfinite(b["loss"]) and b["loss"] < a["loss"]


def test_skip_connect_flag_is_inert_for_uvrgcn():
    """--skip-connect with the uvrgcn encoder: the cell hands prev_h=[] to every layer (src/rrgcn.py:37-38), so scores
    equal those of the same weights without the flag, on the engine path too, and the gate weights get no gradient."""
    R._lib.require_device()
    case = synth.make_case("tiny", 5)
    n, r = case["num_ents"], case["num_rels"]
    models = []
    for skip in (True, False):
        m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, 200, "sub", 3, num_bases=100, num_basis=-1,
                            num_hidden_layers=2, dropout=0.0, self_loop=True, skip_connect=skip, layer_norm=True,
                            input_dropout=0.0, hidden_dropout=0.0, feat_dropout=0.0, entity_prediction=True,
                            relation_prediction=True, use_cuda=True, gpu=0)
        models.append(m)
    sd = synth.fill_state_dict(models[0].state_dict(), 5)
    models[0].load_state_dict(sd)
    models[1].load_state_dict({k: v for k, v in sd.items() if "skip_connect" not in k})
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    triples = torch.from_numpy(case["test"]).to(DEV)
    scores = []
    for m in models:
        m = m.to(DEV).eval()
        assert m._engine_ok()
        _, score, score_rel = m.predict(glist, r, None, triples, True)
        scores.append((score.clone(), score_rel.clone()))
    assert torch.equal(scores[0][0], scores[1][0]) and torch.equal(scores[0][1], scores[1][1])
    m = models[0].train()
    le, lr_, ls = m.get_loss(glist, triples, None, True)
    (0.7 * le + 0.3 * lr_ + ls).backward()
    assert m.rgcn.layers[1].skip_connect_weight.grad is None and m.rgcn.layers[1].skip_connect_bias.grad is None


@pytest.mark.parametrize("which", ["decoder_ob", "rdecoder"])
def test_decoder_forward_in_train_mode_carries_gradients(which):
    """ConvTransE / ConvTransR called directly in train() mode (src/decoder.py:78-100, 29-52 with batch-statistics
    BatchNorm): the (B, N) / (B, 2R) score matrix and the gradients a caller's own loss sends through it."""
    R._lib.require_device()
    case = synth.make_case("tiny", 6)
    n, r = case["num_ents"], case["num_rels"]
    m, sd = _decoder_pair(n, r, 6)
    mod = getattr(m, which).train()
    rng = np.random.default_rng(8)
    d = 200
    emb = rng.standard_normal((n, d))
    rel = rng.standard_normal((2 * r, d)) * 0.3
    all_t = restate.add_inverse(case["test"], r)
    ed, rd = _leaf(emb), _leaf(rel)
    score = mod(ed, rd, torch.from_numpy(all_t).to(DEV), mode="train")
    assert score.requires_grad and score.shape == (len(all_t), n if which == "decoder_ob" else 2 * r)
    w = rng.standard_normal(tuple(score.shape))
    (score * torch.as_tensor(w, dtype=torch.float32, device=DEV)).sum().backward()
    P = {k: v.clone().double() for k, v in sd.items() if v.is_floating_point()}
    names = [k for k in P if k.startswith(which + ".") and "running" not in k and ".bn3." not in k and ".bn_init." not in k
             and not k.endswith(".b")]
    for k in names:
        P[k].requires_grad_(True)
    ec = torch.tensor(emb, dtype=torch.float64, requires_grad=True)
    rc = torch.tensor(rel, dtype=torch.float64, requires_grad=True)
    tt = torch.as_tensor(all_t)
    ea = torch.tanh(ec)
    if which == "decoder_ob":
        sc = restate.conv_tower_train(ea[tt[:, 0]], rc[tt[:, 1]], P, which + ".", {}) @ ea.t()
    else:
        sc = restate.conv_tower_train(ea[tt[:, 0]], ea[tt[:, 2]], P, which + ".", {}) @ rc.t()
    (sc * torch.as_tensor(w)).sum().backward()
    ok, worst = close(score.detach().cpu().numpy(), sc.detach().numpy(), rtol=1e-4)
    assert ok, worst
    mine = [ed.grad, rd.grad] + [dict(m.named_parameters())[k].grad for k in names]
    assert all(g is not None for g in mine)
    _cmp_grads(mine, [ec.grad, rc.grad] + [P[k].grad for k in names], ["d emb", "d rel"] + names)


@pytest.mark.parametrize("ln", [True, False])
def test_model_forward_in_train_mode_carries_gradients(ln):
    """RecurrentRGCN.forward() in train() mode with autograd on (src/rrgcn.py:142-180 is autograd-visible in the
    reference): the history embeddings carry gradients.  With the layer dropout set to 0 the values equal the inference
    engine's, and the gradients of a fixed linear functional of (history_embs[-1], history_embs[0], h_0) equal torch
    autograd over the CPU oracle.  Under torch.no_grad() -- and in eval() mode -- forward() stays the inference path."""
    from tests.helpers import build_model
    case = synth.make_case("small", 5)
    n, r = case["num_ents"], case["num_rels"]
    m, sd = build_model(dict(kind="regcn", layer_norm=ln, seed=5), n, r)
    m = m.to(DEV)
    for layer in m.rgcn.layers:
        layer.dropout.p = 0.0
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    graphs = [restate.build_edges(s, n, r) for s in case["history"]]
    hist_e, _, h0_e, _, _ = m.forward(glist, None, True)
    assert not hist_e[-1].requires_grad
    m.train()
    with torch.no_grad():
        hist_n, _, _, _, _ = m.forward(glist, None, True)
    assert not hist_n[-1].requires_grad and torch.equal(hist_n[-1], hist_e[-1])
    hist, static_emb, h0, gates, degs = m.forward(glist, None, True)
    assert hist[-1].requires_grad and h0.requires_grad and static_emb is None and gates == [] and degs == []
    for a, b in ((hist[-1], hist_e[-1]), (hist[0], hist_e[0]), (h0, h0_e)):
        ok, worst = close(a.detach().cpu().numpy(), b.cpu().numpy(), rtol=1e-4)    # (the tape runs the dense layer form)
        assert ok, worst
    gen = torch.Generator().manual_seed(17)
    w1, w2, w3 = (torch.randn(*t.shape, generator=gen) for t in (hist[-1], hist[0], h0))
    loss = (hist[-1] * w1.to(DEV)).sum() + 0.5 * (hist[0] * w2.to(DEV)).sum() + (h0 * w3.to(DEV)).sum()
    loss.backward()
    P = {k: v.clone().double().requires_grad_(v.is_floating_point()) for k, v in sd.items() if v.is_floating_point()}
    o_hist, o_h0 = restate.regcn_forward(P, graphs, r, layer_norm=ln, dtype=torch.float64)
    o_loss = (o_hist[-1] * w1.double()).sum() + 0.5 * (o_hist[0] * w2.double()).sum() + (o_h0 * w3.double()).sum()
    o_loss.backward()
    np.testing.assert_allclose(float(loss.detach()), float(o_loss.detach()), rtol=1e-5)
    names = ["dynamic_emb", "emb_rel", "time_gate_weight", "time_gate_bias", "relation_cell_1.weight_ih",
             "relation_cell_1.weight_hh", "rgcn.layers.0.weight_neighbor", "rgcn.layers.0.loop_weight",
             "rgcn.layers.1.evolve_loop_weight"]
    params = dict(m.named_parameters())
    mine = [params[k].grad for k in names]
    ref = [P[k].grad.float() for k in names]
    assert all(g is not None for g in mine)
    # whole-recurrence gate (as tests/helpers.compare_train_step): 1.5e-3 of the largest gradient entry, norm-wise 5e-3.
    # Measured: ~2e-6 relative in L2 when no rrelu input sits within rounding of zero, ~5e-4 when ONE mask differs between
    # the 3xTF32 forward and the fp64 oracle (an element of a 500 x 200 layer carries ~1e-3 of the gradient norm).
    tn = float(np.sqrt(sum(float((g_.double() ** 2).sum()) for g_ in ref)))
    for g_, r_, k in zip(mine, ref, names):
        ok, worst = grad_close(g_.detach().cpu().numpy(), r_.numpy(), tn, 1.5e-3)
        l2 = float((g_.detach().cpu().double() - r_.double()).norm()) / max(float(r_.double().norm()), 1e-3 * tn)
        assert ok or l2 <= 5e-3, (k, worst, l2)
