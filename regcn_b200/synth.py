"""Synthetic workload generator (inputs only -- no hot-path arithmetic): used by bench.py, tests/ and
oracle/gen_golden.py.

Deterministic synthetic inputs of the shapes SURVEY.md section 8(d) names: temporal-KG snapshots,
query triples and model parameters, all drawn from numpy's default_rng (bit-stable across machines), so
that the golden fixtures only need to store (config, seed, outputs) and the GPU box can regenerate the
very same inputs without /root/reference.
"""
import math

import numpy as np
import torch

# name -> (num_ents N, num_rels R, triples per snapshot T, history length L, query triples Tq)
SHAPES = {
    "tiny": (64, 6, 40, 3, 24),
    "small": (500, 12, 300, 3, 120),
    # lgcn clamps num_bases to 2R (hyperbolic_layers.py:559-561) and needs it to divide d=200:
    "tiny_l": (64, 5, 40, 3, 24),         # 2R = 10 bases of 20x20
    "small_l": (500, 50, 300, 3, 120),    # 2R = 100 bases of 2x2
    "c1": (7128, 230, 250, 3, 250),       # ICEWS14s-shaped
    "c3": (23033, 256, 1541, 6, 1457),    # ICEWS18-shaped (BASELINE.json headline config)
    "c4": (7691, 240, 5000, 3, 750),      # GDELT-shaped, dense snapshot
    "c4d": (7691, 240, 50000, 3, 750),    # GDELT-shaped, dense stress
}


def _entities(rng, n, size, zipf):
    if not zipf:
        return rng.integers(0, n, size=size, dtype=np.int64)
    # Zipf(alpha=1) over a random permutation of the entity ids: realistic hub skew
    w = 1.0 / np.arange(1, n + 1, dtype=np.float64)
    cdf = np.cumsum(w / w.sum())
    perm = rng.permutation(n)
    return perm[np.minimum(np.searchsorted(cdf, rng.random(size)), n - 1)].astype(np.int64)


def make_snapshot(rng, n, r, t, zipf=True):
    """(t,3) int64 triples (s, r, o); duplicates kept (the reference builds multigraphs)."""
    return np.stack([_entities(rng, n, t, zipf), rng.integers(0, r, size=t, dtype=np.int64),
                     _entities(rng, n, t, zipf)], axis=1)


def make_case(shape="tiny", seed=0, zipf=True):
    n, r, t, hist, tq = SHAPES[shape] if isinstance(shape, str) else shape
    rng = np.random.default_rng(seed)
    history = [make_snapshot(rng, n, r, t, zipf) for _ in range(hist)]
    test = make_snapshot(rng, n, r, tq, zipf)
    return {"num_ents": n, "num_rels": r, "history": history, "test": test}


def make_stream(shape="tiny", seed=0, n_test=4, zipf=True):
    """A sliding-window evaluation stream for the reference's test() loop (src/main.py:33-123): `hist` history
    snapshots of `t` triples, then `n_test` test snapshots of `tq` triples; every evaluated snapshot joins the history
    window afterwards.  Same generator / seeding as make_case (the first hist + 1 snapshots are identical to it)."""
    n, r, t, hist, tq = SHAPES[shape] if isinstance(shape, str) else shape
    rng = np.random.default_rng(seed)
    history = [make_snapshot(rng, n, r, t, zipf) for _ in range(hist)]
    tests = [make_snapshot(rng, n, r, tq, zipf) for _ in range(n_test)]
    return {"num_ents": n, "num_rels": r, "history": history, "tests": tests}


def make_static(num_ents, seed=0, num_static_rels=3, num_words=20, per_entity=2):
    """Synthetic entity-word graph in the layout of the reference's e-w-graph.txt after src/main.py:146-150: (T,3) int64
    triples (entity, static relation, num_ents + word); every word and every static relation occurs at least once.
    Returns (static_triples, num_static_rels, num_words)."""
    rng = np.random.default_rng(50_000 + seed)
    ents = np.repeat(np.arange(num_ents, dtype=np.int64), per_entity)
    rels = rng.integers(0, num_static_rels, size=ents.size, dtype=np.int64)
    words = rng.integers(0, num_words, size=ents.size, dtype=np.int64)
    rels[:num_static_rels] = np.arange(num_static_rels)
    words[:num_words] = np.arange(num_words)
    return np.stack([ents, rels, words + num_ents], axis=1), num_static_rels, num_words


def _scale_for(name, shape):
    leaf = name.split(".")[-1]
    if leaf in ("running_var",):
        return "var"
    if leaf == "num_batches_tracked":
        return "keep"
    if name in ("c", "log_c"):
        return "keep"
    if leaf in ("radius_static", "radius_target"):
        return "radius"
    if leaf in ("score_scale_raw", "score_margin"):
        return "one"
    if leaf == "rel_curvature_raw":
        return "relcurv"
    if ".bn" in name and leaf == "weight":
        return "bnw"
    if leaf in ("bias", "b", "time_gate_bias", "skip_connect_bias", "skip_bias", "rel_bias", "entity_bias",
                "running_mean", "bias_ih", "bias_hh"):
        return 0.05
    if leaf == "global_rot":
        return 1.0
    if leaf == "dynamic_emb":
        return 1.0
    if len(shape) >= 2:
        fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
        if "rot_proj" in name or "trans_proj" in name or "reshape_fc" in name:
            return 0.5 * math.sqrt(2.0 / (fan_in + fan_out))
        return math.sqrt(2.0) * math.sqrt(2.0 / (fan_in + fan_out))
    return 0.1


def fill_state_dict(state_dict, seed):
    """Overwrite every tensor of `state_dict` (name order) with seeded numpy draws.  Returns a new dict.
    Non-trivial BatchNorm statistics / biases are drawn on purpose: zero-initialised tensors would hide bugs."""
    rng = np.random.default_rng(10_000 + seed)
    out = {}
    for name in sorted(state_dict.keys()):
        ref = state_dict[name]
        shape = tuple(ref.shape)
        kind = _scale_for(name, shape)
        if kind == "keep":
            out[name] = ref.clone()
            continue
        if kind == "var":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif kind == "bnw":
            v = rng.uniform(0.8, 1.2, size=shape)
        elif kind == "radius":
            v = rng.uniform(0.6, 2.9, size=shape)
        elif kind == "relcurv":
            v = -4.65 + 0.5 * rng.standard_normal(size=shape)     # softplus ~ 0.006 .. 0.016 around 0.95 c (c = 0.01)
        elif kind == "one":
            v = np.full(shape, 1.0) + rng.uniform(-0.2, 0.2, size=shape)
        else:
            v = rng.standard_normal(size=shape) * kind
        out[name] = torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(shape).to(ref.dtype)
    if "rgcn.rel_emb" in out and "emb_rel" in out:
        out["rgcn.rel_emb"] = out["emb_rel"].clone()  # the reference registers the same Parameter twice
    return out


def answers_of(all_triples_np, num_rels, rel_p=False):
    """The reference's per-snapshot answer dict (rgcn/utils.py:264-283) for the *raw* test triples."""
    d = {}
    for s, r, o in all_triples_np:
        s, r, o = int(s), int(r), int(o)
        if rel_p:
            d.setdefault(s, {}).setdefault(o, set()).add(r)
            d.setdefault(o, {}).setdefault(s, set()).add(r + num_rels)
        else:
            d.setdefault(o, {}).setdefault(r + num_rels, set()).add(s)
            d.setdefault(s, {}).setdefault(r, set()).add(o)
    return d
