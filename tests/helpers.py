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
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(prefix))


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
