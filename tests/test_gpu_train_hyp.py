"""GPU parity of the hyperbolic training step (hyperbolic_uvrgcn + hyperbolic_convtranse): every autograd node of
regcn_b200/train_hyp.py against torch autograd over the CPU oracle, the whole optimisation step against the UNMODIFIED
hyperbolic reference (tests/golden/train_hyp.npz).  Tolerances as in test_gpu_train.py."""
import numpy as np
import pytest
import torch

import regcn_b200 as R
from oracle import restate, synth
from regcn_b200 import optim, train, train_hyp
from tests.helpers import CURV, HYP_TRAIN_CASES, build_hyp_train_model, close, compare_train_step, grad_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
C = 0.01


def _leaf(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32)).to(DEV).requires_grad_(True)


def _cmp(mine, ref, names, rtol=2e-4):
    tn = float(np.sqrt(sum(float((r.double() ** 2).sum()) for r in ref)))
    for m, r, n in zip(mine, ref, names):
        ok, worst = grad_close(m.detach().cpu().numpy(), r.detach().numpy(), tn, rtol)
        assert ok, (n, worst)


def test_radial_row_maps_backward():
    """log_0 / exp_0 / project / tangent-normalise, radius, apply_radius, clamp, leaky tanh, radius_combine, row_dot."""
    R._lib.require_device()
    rng = np.random.default_rng(3)
    n, d = 300, 200
    # rows inside the ball (|x| < 10), some near the boundary, some tiny; tangent rows large enough to hit the projection
    x = rng.standard_normal((n, d)) * rng.uniform(0.01, 0.45, size=(n, 1))   # |x| up to ~6.5 of the radius-10 ball
    x[:5] *= 1e-9
    v = rng.standard_normal((n, d)) * rng.uniform(0.01, 3.0, size=(n, 1))
    go = rng.standard_normal((n, d))
    got = torch.as_tensor(go, dtype=torch.float32, device=DEV)
    for mode, inp, ref_fn in ((train_hyp.LOG0, x, lambda t: restate.log0(t, C)), (train_hyp.EXP0, v, lambda t: restate.exp0(t, C)),
                              (train_hyp.PROJECT, v * 8, lambda t: restate.project(t, C)),
                              # (rows below eps = 1e-6 are outside the tangent normalisation's domain: F.normalize
                              #  divides by the true norm there while log_0 clamps it; entities sit at radius >= 0.5)
                              (train_hyp.TNORM, x[5:], lambda t: restate.exp0(restate.normalize_rows(restate.log0(t, C)), C))):
        xd = _leaf(inp)
        y = train_hyp.radial(xd, mode, C)
        y.backward(got[:len(inp)])
        xc = torch.tensor(inp, dtype=torch.float64, requires_grad=True)
        yc = ref_fn(xc)
        yc.backward(torch.as_tensor(go[:len(inp)]))
        ok, worst = close(y.detach().cpu().numpy(), yc.detach().numpy())
        assert ok, (mode, worst)
        _cmp([xd.grad], [xc.grad], [f"radial mode {mode}"])
    # radius / apply_radius
    r = rng.uniform(0.2, 12.0, size=n)                      # some beyond 1/sqrt(c): clamped
    xd, rd = _leaf(x), _leaf(r)
    y = train_hyp.apply_radius(xd, rd, C)
    rho = train_hyp.radius(xd)
    (y * got).sum().backward(retain_graph=True)
    (rho * torch.arange(n, device=DEV)).sum().backward()
    xc = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    rc = torch.tensor(r, dtype=torch.float64, requires_grad=True)
    yc = restate.apply_radius(xc, rc, C)
    ((yc * torch.as_tensor(go)).sum() + (restate.get_radius(xc) * torch.arange(n)).sum()).backward()
    ok, worst = close(y.detach().cpu().numpy(), yc.detach().numpy())
    assert ok, worst
    _cmp([xd.grad, rd.grad], [xc.grad, rc.grad], ["apply_radius dx (+ radius)", "apply_radius dr"])
    # elementwise
    for op, ref_fn in ((0, lambda t: t.clamp(-1.5, 1.5)), (1, lambda t: 0.9 * torch.tanh(t) + 0.1 * t)):
        xd = _leaf(v)
        train_hyp.eltwise(xd, op, 1.5).backward(got)
        xc = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        ref_fn(xc).backward(torch.as_tensor(go))
        _cmp([xd.grad], [xc.grad], [f"eltwise {op}"])
    # radius_combine + row_dot (TemporalRadiusEvolution scalars)
    raw = rng.uniform(0.2, 3.5, size=n)
    w, b = rng.standard_normal((1, d)) * 0.05, rng.standard_normal(1) * 0.05
    gv = rng.standard_normal(n)
    rawd, xd, wd, bd = _leaf(raw), _leaf(x), _leaf(w), _leaf(b)
    t = train_hyp.radial(xd, train_hyp.LOG0, C)
    delta = train_hyp.row_dot(t, wd, bd)
    out = train_hyp._RadiusCombine.apply(rawd, train_hyp.radius(xd), delta, 0.5, 3.0, C, 0.7, 0.1)
    out.backward(torch.as_tensor(gv, dtype=torch.float32, device=DEV))
    rawc, xc, wc, bc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (raw, x, w, b))
    tc = restate.log0(xc, C)
    dc = (tc @ wc.t()).squeeze(-1) + bc
    rs = restate.static_radius(rawc, C, 0.5, 3.0)
    outc = 0.7 * rs + 0.3 * restate.get_radius(xc) + dc.clamp(-0.1, 0.1)
    outc.backward(torch.as_tensor(gv))
    ok, worst = close(out.detach().cpu().numpy(), outc.detach().numpy())
    assert ok, worst
    _cmp([rawd.grad, xd.grad, wd.grad, bd.grad], [rawc.grad, xc.grad, wc.grad, bc.grad], ["draw", "dx", "dw", "db"])


def test_hyperbolic_union_layer_backward():
    """One HyperbolicUnionRGCNLayer (radius-weighted aggregate incl. the gradient w.r.t. the radii) against the oracle."""
    R._lib.require_device()
    case = synth.make_case("small", 6)
    n, r = case["num_ents"], case["num_rels"]
    g = R.build_sub_graph(n, r, case["history"][0], True, 0)
    og = restate.build_edges(case["history"][0], n, r)
    rng = np.random.default_rng(8)
    d = 200
    h = rng.standard_normal((n, d)) * rng.uniform(0.02, 0.25, size=(n, 1))
    rel = rng.standard_normal((2 * r, d)) * 0.2
    W = [rng.standard_normal((d, d)) * 0.1 for _ in range(3)]
    go = rng.standard_normal((n, d))
    layer = R.HyperbolicUnionRGCNLayer(d, d, 2 * r, c=C, activation=torch.nn.functional.rrelu, self_loop=True,
                                       dropout=0.0, radius_msg_gamma=0.6).to(DEV)
    with torch.no_grad():
        for p, w in zip((layer.weight_neighbor, layer.loop_weight, layer.evolve_loop_weight), W):
            p.copy_(torch.as_tensor(w, dtype=torch.float32))
    hd, rd = _leaf(h), _leaf(rel)
    out = train_hyp.hyp_union_layer(layer, g, hd, rd, C, True)
    out.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    hc, rc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (h, rel))
    Wc = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in W]
    outc = restate.hyp_union_layer(hc, rc, og, Wc[0], Wc[1], Wc[2], C, 0.6)
    outc.backward(torch.as_tensor(go))
    ok, worst = close(out.detach().cpu().numpy(), outc.detach().numpy())
    assert ok, worst
    _cmp([hd.grad, rd.grad, layer.weight_neighbor.grad, layer.loop_weight.grad, layer.evolve_loop_weight.grad],
         [hc.grad, rc.grad] + [w.grad for w in Wc], ["dh", "drel", "dW_n", "dW_loop", "dW_evolve"])


@pytest.mark.parametrize("d,nb", [(200, 100), (200, 50), (200, 10), (200, 1), (64, 16), (64, 4), (128, 2)])
def test_lorentz_layer_backward_any_block_size(d, nb):
    """One LorentzRGCNLayer (lgcn) in train() mode against torch autograd over the fp64 oracle: relation blocks of 2x2 (the
    chunk-local kernels) and of 4x4 ... dxd (the generic kernels; num_bases clamped to 2R, hyperbolic_layers.py:559-561),
    both register widths (d <= 128 / d > 128)."""
    R._lib.require_device()
    case = synth.make_case("small_l", 6)
    n, r = case["num_ents"], case["num_rels"]
    g = R.build_sub_graph(n, r, case["history"][1], True, 0)
    og = restate.build_edges(case["history"][1], n, r)
    rng = np.random.default_rng(100 * d + nb)
    sb = d // nb
    h = rng.standard_normal((n, d)) * rng.uniform(0.02, 0.25, size=(n, 1))
    rel = rng.standard_normal((2 * r, d)) * 0.2
    W = rng.standard_normal((2 * r, nb * sb * sb)) * (0.5 / np.sqrt(sb))
    wl, we = rng.standard_normal((d, d)) * 0.1, rng.standard_normal((d, d)) * 0.1
    go = rng.standard_normal((n, d))
    layer = R.LorentzRGCNLayer(d, d, 2 * r, num_bases=nb, c=C, activation=torch.nn.functional.rrelu, self_loop=True,
                               dropout=0.0).to(DEV)
    assert layer.num_bases == nb
    with torch.no_grad():
        for p_, w in zip((layer.weight, layer.loop_weight, layer.evolve_loop_weight), (W, wl, we)):
            p_.copy_(torch.as_tensor(w, dtype=torch.float32))
    hd, rd = _leaf(h), _leaf(rel)
    out = train_hyp.lorentz_layer(layer, g, hd, rd, C, True)
    out.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    hc, rc, Wc, wlc, wec = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (h, rel, W, wl, we))
    outc = restate.lorentz_layer(hc, rc, og, Wc, wlc, wec, C, nb)
    outc.backward(torch.as_tensor(go))
    ok, worst = close(out.detach().cpu().numpy(), outc.detach().numpy())
    assert ok, worst
    _cmp([hd.grad, rd.grad, layer.weight.grad, layer.loop_weight.grad, layer.evolve_loop_weight.grad],
         [hc.grad, rc.grad, Wc.grad, wlc.grad, wec.grad], ["dh", "drel", "dW_blocks", "dW_loop", "dW_evolve"])
    # bit-reproducible (fixed summation order in every kernel of the backward)
    g1 = layer.weight.grad.clone()
    layer.weight.grad = None
    hd2, rd2 = _leaf(h), _leaf(rel)
    train_hyp.lorentz_layer(layer, g, hd2, rd2, C, True).backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    assert torch.equal(g1, layer.weight.grad) and torch.equal(hd.grad, hd2.grad)


@pytest.mark.parametrize("name", sorted(HYP_TRAIN_CASES))
def test_hyperbolic_train_step_matches_reference(name):
    """One optimisation step of HyperbolicRecurrentRGCN (get_loss in train() mode -> backward -> clip -> Adam) against the
    UNMODIFIED reference: four losses, gradient norm, all 39 parameter gradients, updated values."""
    import os
    from tests.helpers import GOLDEN
    R._lib.require_device()
    z = np.load(os.path.join(GOLDEN, "train_hyp.npz"))
    cfg = HYP_TRAIN_CASES[name]
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    m, _ = build_hyp_train_model(cfg, n, r)
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    sg = None
    if cfg.get("static") or cfg.get("skip_connect"):
        # --add-static-graph (hyperbolic_src/hyperbolic_model.py:762-771,1039-1064) / --skip-connect: eval outputs first
        if cfg.get("static"):
            st, n_srel, n_words = synth.make_static(n, cfg["seed"])
            sg = R.build_sub_graph(n + n_words, n_srel, st, True, 0)
        m = m.to(DEV).eval()
        triples = torch.from_numpy(case["test"]).to(DEV)
        # static initial table / inert skip flag through the one-call evolve engine; the live Lorentz skip gate
        # (hyperbolic_src/hyperbolic_layers.py:657-678) on the per-layer path
        assert m._engine_ok() == (not (cfg["encoder"] == "lgcn" and cfg.get("skip_connect")))
        _, score, score_rel = m.predict(glist, r, sg, triples, True)
        hist, static_emb, _, _, _ = m.forward(glist, sg, True)
        pairs = [(hist[-1], "hist_last"), (score, "score"), (score_rel, "score_rel")]
        for mine, key in pairs + ([(static_emb, "static_emb")] if sg is not None else []):
            ok, worst = close(mine.cpu().numpy(), z[f"{name}.{key}"], rtol=1e-4)
            assert ok, (key, worst)
        ev = m.get_loss(glist, triples, sg, True)
        np.testing.assert_allclose([float(x.reshape(-1)[0]) for x in ev], z[f"{name}.eval_losses"], rtol=1e-4)
    m = m.to(DEV).train()
    opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    le, lr_, ls, lrad = m.get_loss(glist, torch.from_numpy(case["test"]).to(DEV), sg, True)
    (0.7 * le + 0.3 * lr_ + ls + lrad).backward()
    named = {k: p for k, p in m.named_parameters() if p.grad is not None}
    grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in named.items()}
    optim.clip_grad_norm_(opt, 1.0)
    opt.step()
    params = {k: p.detach().cpu().numpy().copy() for k, p in named.items()}
    losses = tuple(float(x.detach().reshape(-1)[0]) for x in (le, lr_, ls, lrad))
    compare_train_step(z, name, 0, losses, float(opt.total_norm), grads, params, rtol=1.5e-3)
    # and the trained model still evaluates (operand caches follow the version counters)
    m.eval()
    _, score, _ = m.predict(glist, r, sg, torch.from_numpy(case["test"]).to(DEV), True)
    assert torch.isfinite(score).all()


def test_distance_decoder_nodes_backward():
    """mobius_add, gather, elementwise product and the all-candidate hyperbolic-distance cross entropy (dq, dE, dbias,
    dscale, dmargin through the <q,e>, |q|^2, |e|^2 form) against autograd on the oracle."""
    R._lib.require_device()
    rng = np.random.default_rng(21)
    B, N, d = 96, 301, 200
    q = rng.standard_normal((B, d)) * rng.uniform(0.02, 0.3, size=(B, 1))
    E = rng.standard_normal((N, d)) * rng.uniform(0.02, 0.3, size=(N, 1))
    bias = rng.standard_normal(N) * 0.1
    trip = np.stack([rng.integers(0, N, B), rng.integers(0, 5, B), rng.integers(0, N, B)], 1).astype(np.int64)
    qd, Ed, bd = _leaf(q), _leaf(E), _leaf(bias)
    sc, mg = _leaf(np.array(1.3)), _leaf(np.array(0.8))
    loss = train_hyp._HypDistCE.apply(qd, Ed, bd, sc, mg, torch.from_numpy(trip).to(DEV), 2, C)
    (loss * 1.5).backward()
    qc, Ec, bc = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (q, E, bias))
    scc, mgc = torch.tensor(1.3, dtype=torch.float64, requires_grad=True), torch.tensor(0.8, dtype=torch.float64, requires_grad=True)
    S = restate.hyp_dist_scores(qc, Ec, bc, C, scc, mgc)
    lc = restate.cross_entropy(S, trip[:, 2])
    (lc * 1.5).backward()
    assert abs(float(loss.detach()) - float(lc.detach())) <= 1e-4 * max(1.0, abs(float(lc.detach())))
    _cmp([qd.grad, Ed.grad, bd.grad, sc.grad.view(1), mg.grad.view(1)], [qc.grad, Ec.grad, bc.grad, scc.grad.view(1), mgc.grad.view(1)],
         ["dq", "dE", "dbias", "dscale", "dmargin"])
    # mobius + product + gather
    x, y = q, rng.standard_normal((B, d)) * 0.1
    go = rng.standard_normal((B, d))
    idx = rng.integers(0, N, B).astype(np.int32)
    xd, yd, Ed = _leaf(x), _leaf(y), _leaf(E)
    z = train_hyp._Mobius.apply(train_hyp._Mul.apply(xd, train_hyp._GatherRows.apply(Ed, torch.from_numpy(idx).to(DEV))), yd, C)
    z = train_hyp.radial(z, train_hyp.PROJECT, C)
    z.backward(torch.as_tensor(go, dtype=torch.float32, device=DEV))
    xc, yc, Ec = (torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (x, y, E))
    zc = restate.mobius_add(xc * Ec[torch.as_tensor(idx, dtype=torch.long)], yc, C)
    zc.backward(torch.as_tensor(go))
    ok, worst = close(z.detach().cpu().numpy(), zc.detach().numpy())
    assert ok, worst
    _cmp([xd.grad, yd.grad, Ed.grad], [xc.grad, yc.grad, Ec.grad], ["mobius dx", "mobius dy", "gather dE"])


@pytest.mark.parametrize("decoder", ["murp", "roth", "atth"])
def test_decoder_loss_methods_match_oracle(decoder):
    """The decoders' own training heads -- Hyperbolic{MuRP,RotH,AttH}.loss and ...Rel.loss
    (hyperbolic_src/hyperbolic_decoder.py:781,897,1101,1249,1464,1641) -- called directly on an entity / relation table:
    the scalar and its gradients w.r.t. the entity table and every decoder parameter against autograd on the oracle."""
    R._lib.require_device()
    cfg = dict(kind="hyp", shape="tiny", seed=31, encoder="hyperbolic_uvrgcn", decoder=decoder, layer_norm=False,
               gamma=0.15, entity_bias=True)
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    m, sd = build_hyp_train_model(cfg, n, r)
    m = m.to(DEV).train()
    rng = np.random.default_rng(5)
    emb = rng.standard_normal((n, 200)) * rng.uniform(0.05, 0.4, size=(n, 1))
    rel = rng.standard_normal((2 * r, 200)) * 0.3
    t = case["test"]
    inv = t[:, ::-1].copy()
    inv[:, 1] += r
    all_t = np.concatenate([t, inv]).astype(np.int64)
    ed, rd = _leaf(emb), _leaf(rel)
    le = m.decoder_ob.loss(ed, rd, torch.from_numpy(all_t).to(DEV))
    lr_ = m.rdecoder.loss(ed, rd, torch.from_numpy(all_t).to(DEV))
    assert le.dim() == 0 and lr_.dim() == 0
    (le + 0.5 * lr_).backward()
    P = {k: v.detach().double().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    ec = torch.tensor(emb, dtype=torch.float64, requires_grad=True)
    rc = torch.tensor(rel, dtype=torch.float64, requires_grad=True)
    ent_fn, rel_fn = {"murp": (restate.murp_scores, restate.murprel_scores),
                      "roth": (restate.roth_scores, restate.rothrel_scores),
                      "atth": (restate.atth_scores, restate.atthrel_scores)}[decoder]
    tt = torch.from_numpy(all_t)
    sc = ent_fn(P, ec, rc, all_t, CURV)[0] - P["decoder_ob.entity_bias"][tt[:, 0]].unsqueeze(1)
    lec = restate.cross_entropy(sc, tt[:, 2])
    lrc = restate.cross_entropy(rel_fn(P, ec, rc, all_t, CURV)[0], tt[:, 1])
    (lec + 0.5 * lrc).backward()
    for mine, ref in ((le, lec), (lr_, lrc)):
        assert abs(float(mine.detach()) - float(ref.detach())) <= 1e-4 * max(1.0, abs(float(ref.detach())))
    names = [k for k, p in m.named_parameters() if k.startswith(("decoder_ob.", "rdecoder.")) and p.grad is not None]
    assert names and all(P[k].grad is not None for k in names)
    params = dict(m.named_parameters())
    _cmp([ed.grad, rd.grad] + [params[k].grad.reshape(P[k].shape) for k in names],
         [ec.grad, rc.grad] + [P[k].grad for k in names], ["dE", "drel"] + names)


@pytest.mark.parametrize("which", ["decoder_ob", "rdecoder"])
def test_hyperbolic_convtrans_forward_in_train_mode(which):
    """HyperbolicConvTransE / ConvTransR called directly in train() mode (hyperbolic_src/hyperbolic_decoder.py:360-413,
    464-510): scores with the candidate bias, and the gradients of a caller's loss through them."""
    R._lib.require_device()
    cfg = dict(kind="hyp", shape="tiny", seed=33, encoder="hyperbolic_uvrgcn", decoder="hyperbolic_convtranse",
               layer_norm=False, gamma=0.15)
    case = synth.make_case(cfg["shape"], cfg["seed"])
    n, r = case["num_ents"], case["num_rels"]
    m, sd = build_hyp_train_model(cfg, n, r)
    m = m.to(DEV).train()
    mod = getattr(m, which)
    rng = np.random.default_rng(9)
    emb = rng.standard_normal((n, 200)) * rng.uniform(0.05, 0.4, size=(n, 1))
    rel = rng.standard_normal((2 * r, 200)) * 0.3
    all_t = restate.add_inverse(case["test"], r)
    ed, rd = _leaf(emb), _leaf(rel)
    score = mod(ed, rd, torch.from_numpy(all_t).to(DEV), mode="train")
    assert score.requires_grad
    w = rng.standard_normal(tuple(score.shape))
    (score * torch.as_tensor(w, dtype=torch.float32, device=DEV)).sum().backward()
    P = {k: v.clone().double() for k, v in sd.items() if v.is_floating_point()}
    names = [k for k in P if k.startswith(which + ".") and "running" not in k and ".bn3." not in k and ".bn_init." not in k]
    for k in names:
        P[k].requires_grad_(True)
    ec = torch.tensor(emb, dtype=torch.float64, requires_grad=True)
    rc = torch.tensor(rel, dtype=torch.float64, requires_grad=True)
    tt = torch.as_tensor(all_t)
    et = restate.log0(ec, CURV)
    et = 0.9 * torch.tanh(et) + 0.1 * et
    if which == "decoder_ob":
        sc = restate.conv_tower_train(et[tt[:, 0]], rc[tt[:, 1]], P, which + ".", {}) @ et.t() + P[which + ".b"]
    else:
        sc = restate.conv_tower_train(et[tt[:, 0]], et[tt[:, 2]], P, which + ".", {}) @ rc.t() + P[which + ".b"]
    (sc * torch.as_tensor(w)).sum().backward()
    ok, worst = close(score.detach().cpu().numpy(), sc.detach().numpy(), rtol=1e-4)
    assert ok, worst
    params = dict(m.named_parameters())
    mine = [ed.grad, rd.grad] + [params[k].grad for k in names]
    assert all(g is not None for g in mine), [k for k, g in zip(["dE", "drel"] + names, mine) if g is None]
    _cmp(mine, [ec.grad, rc.grad] + [P[k].grad for k in names], ["dE", "drel"] + names)


def test_hyperbolic_forward_in_train_mode_carries_gradients():
    """HyperbolicRecurrentRGCN.forward() in train() mode with autograd on returns history embeddings with gradients (the
    autograd nodes get_loss() trains through); with dropout 0 the values equal the inference engine's."""
    import regcn_b200 as R
    from oracle import synth
    from tests.helpers import build_model, close
    case = synth.make_case("small", 4)
    n, r = case["num_ents"], case["num_rels"]
    m, _ = build_model(dict(kind="hyp", layer_norm=False, seed=4, encoder="hyperbolic_uvrgcn", decoder="roth", gamma=0.15), n, r)
    m = m.to("cuda")
    for layer in m.rgcn.layers:
        if getattr(layer, "dropout", None) is not None:
            layer.dropout.p = 0.0
    glist = [R.build_sub_graph(n, r, s, True, 0) for s in case["history"]]
    hist_e, _, h0_e, _, _ = m.forward(glist, None, True)
    m.train()
    hist, _, h0, _, _ = m.forward(glist, None, True)
    assert hist[-1].requires_grad and h0.requires_grad
    ok, worst = close(hist[-1].detach().cpu().numpy(), hist_e[-1].cpu().numpy(), rtol=5e-5)
    assert ok, worst
    (hist[-1].square().sum() + h0.square().sum()).backward()
    g = m.dynamic_emb.grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0
    assert m.time_gate_weight.grad is not None and float(m.time_gate_weight.grad.abs().max()) > 0
