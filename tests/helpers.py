"""Shared test helpers: rebuild a golden case's inputs (oracle/synth.py) and the model that consumes them."""
import json
import os

import numpy as np
import torch

from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
H_DIM, N_BASES, N_LAYERS, CURV = 200, 100, 2, 0.01


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN)
                  if f.endswith(".npz") and f.startswith(prefix) and not f.startswith(("train_", "aux_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["config"]))
    return cfg, z


def build_model(cfg, n, r, h_dim=H_DIM):
    """The product module for a golden config, parameters filled by the same seeded recipe the reference got."""
    import regcn_b200 as R
    if cfg["kind"] == "regcn":
        m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, 0, 0, h_dim, "sub", 3, num_bases=N_BASES, num_basis=-1,
                            num_hidden_layers=N_LAYERS, dropout=0.2, self_loop=True, skip_connect=False,
                            layer_norm=cfg["layer_norm"], input_dropout=0.2, hidden_dropout=0.2, feat_dropout=0.2,
                            entity_prediction=True, relation_prediction=True, use_cuda=True, gpu=0)
    else:
        m = R.HyperbolicRecurrentRGCN(cfg["decoder"], cfg["encoder"], n, r, 0, 0, h_dim, "sub", 3, num_bases=N_BASES,
                                      num_hidden_layers=N_LAYERS, dropout=0.2, c=CURV, self_loop=True,
                                      skip_connect=False, layer_norm=cfg["layer_norm"], input_dropout=0.2,
                                      hidden_dropout=0.2, feat_dropout=0.2, entity_prediction=True,
                                      relation_prediction=True, use_cuda=True, gpu=0,
                                      radius_msg_gamma=cfg["gamma"], hyp_init_scale=1e-3,
                                      use_entity_euclidean_bias=cfg.get("entity_bias", False),
                                      use_relation_specific_curvature=cfg.get("rel_curvature", False))
    sd = synth.fill_state_dict(m.state_dict(), cfg["seed"])
    m.load_state_dict(sd)
    m.eval()
    return m, sd


def close(a, b, rtol=1e-4, atol_scale=1.0):
    """The parity gate of SURVEY.md 8(d): |a-b| <= rtol * max(1, |b|) elementwise.  Returns (ok, worst ratio)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    tol = rtol * np.maximum(atol_scale, np.abs(b))
    ratio = np.abs(a - b) / tol
    return bool(np.all(ratio <= 1.0)), float(ratio.max()) if ratio.size else 0.0


def grad_close(g, g_ref, total_norm, rtol=2e-4, kink_floor=0.0):
    """Gradient parity gate: |g - g_ref| <= rtol * max(max|g_ref|, 1e-3 * total gradient norm) elementwise.  The floor
    covers gradients that are analytically zero (e.g. a bias in front of a BatchNorm) and differ only by rounding noise.
    Returns (ok, worst ratio)."""
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    g_ref = np.asarray(g_ref, dtype=np.float64).reshape(-1)
    tol = rtol * max(float(np.max(np.abs(g_ref))) if g_ref.size else 0.0, 1e-3 * float(total_norm))
    # kink_floor (whole-model steps only): a ReLU input within rounding of zero flips its mask between two correct fp32
    # implementations and moves a per-channel sum (BatchNorm / conv parameter gradients) by one whole term
    tol = max(tol, kink_floor * float(total_norm))
    worst = float(np.max(np.abs(g - g_ref)) / tol) if g.size else 0.0
    return worst <= 1.0, worst


def sample_of(x, n=1024):
    """The fixed-stride sample oracle/gen_golden.py stores for big tensors."""
    flat = np.asarray(x, dtype=np.float32).reshape(-1)
    step = max(1, flat.size // n)
    return flat[::step][:n]


def compare_train_step(z, name, step, losses, grad_norm, grads, params, lr=1e-3, loss_rtol=2e-5, rtol=2e-4,
                       kink_floor=2e-5):
    """One optimisation step against tests/golden/train_regcn.npz (the UNMODIFIED reference, oracle/gen_golden.py
    --train).  grads / params: {parameter name: full numpy array} (raw, un-clipped gradients; values after the update).

    Step 0 is gated tightly: losses, total gradient norm, every gradient (grad_close), every updated value.  Adam's
    first update is -lr * g / (|g| + eps): an element whose gradient is rounding noise moves by a full +-lr in a
    direction no two fp32 implementations agree on, so such elements are pinned to 2.1 lr only -- and from the second
    step on the two runs sit at (slightly) different points: there the gate is the losses, the total norm, each
    gradient's norm and a step-sized bound on the values."""
    tn = float(z[f"{name}.s{step}.grad_norm"])
    np.testing.assert_allclose(losses, z[f"{name}.s{step}.losses"][:len(losses)], rtol=loss_rtol if step == 0 else 5e-4)
    np.testing.assert_allclose(grad_norm, tn, rtol=2e-4 if step == 0 else 2e-3)
    keys = sorted(k[len(f"{name}.s{step}.g."):] for k in z.files if k.startswith(f"{name}.s{step}.g."))
    assert keys == sorted(grads), (set(keys) ^ set(grads))
    worst_all = 0.0
    for k in keys:
        g_ref = z[f"{name}.s{step}.g.{k}"].astype(np.float64)
        p_ref = z[f"{name}.s{step}.p.{k}"].astype(np.float64)
        g = sample_of(grads[k]).astype(np.float64)
        p = sample_of(params[k]).astype(np.float64)
        gn_ref = float(z[f"{name}.s{step}.gn.{k}"])
        gn = float(np.linalg.norm(np.asarray(grads[k], dtype=np.float64)))
        if step == 0:
            ok, worst = grad_close(g, g_ref, tn, rtol, kink_floor)
            if not ok:
                # several mask flips can land in one small tensor (a conv / BatchNorm parameter sums B*d terms per
                # element): fall back to the norm-wise gate, which a wiring error cannot pass
                l2 = float(np.linalg.norm(g - g_ref)) / max(float(np.linalg.norm(g_ref)), 1e-3 * tn)
                assert l2 <= 5e-3, (k, "gradient", worst, l2)
            worst_all = max(worst_all, worst)
            assert abs(gn - gn_ref) <= max(10 * rtol * gn_ref, 4 * kink_floor * tn) + 1e-6 * tn, (k, gn, gn_ref)
            # an element whose reference gradient is within the gradient gate of zero may carry either sign
            floor = max(max(rtol, 1e-3) * max(float(np.max(np.abs(g_ref))), 1e-3 * tn), kink_floor * tn)
            atol = np.where(np.abs(g_ref) < floor, 2.1 * lr, 2e-5)
            assert np.all(np.abs(p - p_ref) <= atol + 1e-4 * np.abs(p_ref)), (k, "value", float(np.max(np.abs(p - p_ref))))
        else:
            assert abs(gn - gn_ref) <= 3e-2 * gn_ref + 1e-4 * tn, (k, gn, gn_ref)
            assert np.all(np.abs(p - p_ref) <= 2.1 * lr * (step + 1)), (k, "value", float(np.max(np.abs(p - p_ref))))
    return worst_all


STATIC_CASES = {"static_tiny_s0": dict(shape="tiny", seed=0, layer_norm=True, discount=1, angle=10, weight=0.5),
                "static_tiny_s1_noln": dict(shape="tiny", seed=1, layer_norm=False, discount=0, angle=10, weight=1.0),
                "static_small_s2": dict(shape="small", seed=2, layer_norm=True, discount=1, angle=10, weight=0.5)}


def build_static_model(cfg, n, r, n_srel, n_words, dropout=0.0):
    """Product RecurrentRGCN with the static-graph constraint on (src/main.py --add-static-graph)."""
    import regcn_b200 as R
    m = R.RecurrentRGCN("convtranse", "uvrgcn", n, r, n_srel, n_words, H_DIM, "sub", 3, num_bases=N_BASES, num_basis=-1,
                        num_hidden_layers=N_LAYERS, dropout=dropout, self_loop=True, skip_connect=False,
                        layer_norm=cfg["layer_norm"], input_dropout=dropout, hidden_dropout=dropout,
                        feat_dropout=dropout, weight=cfg["weight"], discount=cfg["discount"], angle=cfg["angle"],
                        use_static=True, entity_prediction=True, relation_prediction=True, use_cuda=True, gpu=0)
    sd = synth.fill_state_dict(m.state_dict(), cfg["seed"])
    m.load_state_dict(sd)
    return m, sd


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


def build_hyp_train_model(cfg, n, r, dropout=0.0):
    import regcn_b200 as R
    st_cfg, n_srel, n_words = cfg.get("static"), 0, 0
    if st_cfg:
        _, n_srel, n_words = synth.make_static(n, cfg["seed"])
    m = R.HyperbolicRecurrentRGCN(cfg["decoder"], cfg["encoder"], n, r, n_srel, n_words, H_DIM, "sub", 3,
                                  use_static=bool(st_cfg), **(st_cfg or {}), num_bases=N_BASES,
                                  num_hidden_layers=N_LAYERS, dropout=dropout, c=CURV, self_loop=True,
                                  skip_connect=cfg.get("skip_connect", False),
                                  layer_norm=cfg["layer_norm"], input_dropout=dropout, hidden_dropout=dropout,
                                  feat_dropout=dropout, entity_prediction=True, relation_prediction=True, use_cuda=True,
                                  gpu=0, radius_msg_gamma=cfg["gamma"], hyp_init_scale=1e-3,
                                  use_entity_euclidean_bias=cfg.get("entity_bias", False),
                                  use_relation_specific_curvature=cfg.get("rel_curvature", False))
    sd = synth.fill_state_dict(m.state_dict(), cfg["seed"])
    m.load_state_dict(sd)
    return m, sd


def hyp_rgcn_layer_cases():
    """Inputs of tests/golden/aux_layer_hyp_rgcn.npz (oracle/gen_golden.py --layers): yields (k, variant, case, h, prev, sd)."""
    import json
    from oracle import synth
    z = np.load(os.path.join(GOLDEN, "aux_layer_hyp_rgcn.npz"))
    k = 0
    while f"v{k}_config" in z.files:
        v = json.loads(str(z[f"v{k}_config"]))
        case = synth.make_case(v["shape"], 40 + k)
        n = case["num_ents"]
        rng = np.random.default_rng(900 + k)
        h = rng.standard_normal((n, 200)).astype(np.float32)
        h = h / np.linalg.norm(h, axis=1, keepdims=True) * rng.uniform(0.2, 6.0, size=(n, 1)).astype(np.float32)
        prev = rng.standard_normal((n, 200)).astype(np.float32)
        prev = prev / np.linalg.norm(prev, axis=1, keepdims=True) * rng.uniform(0.2, 6.0, size=(n, 1)).astype(np.float32)
        yield k, v, case, h, prev, z[f"v{k}_out"]
        k += 1
