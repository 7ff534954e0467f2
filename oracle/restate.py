"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
package regcn_b200/ never does.  Every function cites the reference file:line it follows (paths relative
to the reference root).  Integer work (edge index, relation sets, ranks) is numpy and exact; floating
point work is plain torch-on-CPU tensor algebra in a caller-chosen dtype (float32 to mimic the reference,
float64 as the "truth" oracle that decides whether a kernel-vs-reference gap is kernel error or fp32 noise).

Pinning: the reference ships no tests or golden vectors.  This restatement is
pinned instead against outputs of the *reference itself*, executed in the build container under
oracle/fake_dgl.py; those outputs are committed as tests/golden/*.npz by oracle/gen_golden.py and checked
by tests/test_oracle_golden.py.
"""
import math

import numpy as np
import torch

EPS = 1e-6                      # HyperbolicOps.EPS, hyperbolic_ops.py:28
RRELU_SLOPE = (1.0 / 8 + 1.0 / 3) / 2   # F.rrelu in eval mode (rgcn/layers.py:250-251 call it functionally)
FILTER_SCORE = -10000000        # rgcn/utils.py:60,74


# =====================================================================================
# Integer part: edge index and relation->entity sets
# =====================================================================================
def build_edges(triples, num_nodes, num_rels):
    """rgcn/utils.py:116-125 + comp_deg_norm :110-114."""
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    s, r, o = triples[:, 0], triples[:, 1], triples[:, 2]
    src = np.concatenate((s, o))
    dst = np.concatenate((o, s))
    etype = np.concatenate((r, r + num_rels))
    indeg = np.bincount(dst, minlength=num_nodes).astype(np.int64)
    deg = indeg.astype(np.float32)
    deg[deg == 0] = 1
    norm = (np.float32(1.0) / deg).astype(np.float32)
    return {"src": src, "dst": dst, "etype": etype, "indeg": indeg, "norm": norm, "num_nodes": num_nodes,
            "num_rels": num_rels, "triples": triples}


def csr_by_dst(g):
    """CSR-by-destination, stable in edge id (the order the CUDA index must reproduce bit-exactly)."""
    eperm = np.argsort(g["dst"], kind="stable")
    rowptr = np.zeros(g["num_nodes"] + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(g["indeg"])
    return rowptr, eperm, g["src"][eperm], g["etype"][eperm]


def r2e(triples, num_rels):
    """rgcn/utils.py:78-97 as a CSR over base relations r < R (r and r+R share the set :85-89); entities of
    a relation are returned sorted (the reference's python-set order is arbitrary and only feeds a mean)."""
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    sets = [set() for _ in range(num_rels)]
    for s, r, o in triples:
        sets[r].add(int(s))
        sets[r].add(int(o))
    rel_rowptr = np.zeros(num_rels + 1, dtype=np.int64)
    ents = []
    for r in range(num_rels):
        e = sorted(sets[r])
        ents.extend(e)
        rel_rowptr[r + 1] = rel_rowptr[r] + len(e)
    return rel_rowptr, np.asarray(ents, dtype=np.int64)


# =====================================================================================
# Euclidean RE-GCN evolution
# =====================================================================================
def normalize_rows(x):
    """F.normalize (src/rrgcn.py:154,170,176,190): x / max(|x|_2, 1e-12)."""
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def rrelu(x):
    return torch.where(x >= 0, x, x * RRELU_SLOPE)


def rel_mean_pool(h, rel_rowptr, rel_ents, num_rels):
    """src/rrgcn.py:161-166: x_input[r] = mean(h[ents(r)]) for present r (and r+R), zeros otherwise."""
    out = torch.zeros(2 * num_rels, h.shape[1], dtype=h.dtype)
    for r in range(num_rels):
        b, e = int(rel_rowptr[r]), int(rel_rowptr[r + 1])
        if e > b:
            m = h[torch.as_tensor(rel_ents[b:e])].mean(dim=0)
            out[r] = m
            out[r + num_rels] = m
    return out


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """nn.GRUCell (src/rrgcn.py:133): gate order r, z, n."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    d = h.shape[1]
    r = torch.sigmoid(gi[:, :d] + gh[:, :d])
    z = torch.sigmoid(gi[:, d:2 * d] + gh[:, d:2 * d])
    n = torch.tanh(gi[:, 2 * d:] + r * gh[:, 2 * d:])
    return (h - n) * z + n


def scatter_sum(msg, dst, num_nodes):
    """DGL update_all(fn.sum): sum of in-edge messages, exact zeros for in-degree 0."""
    out = torch.zeros(num_nodes, msg.shape[1], dtype=msg.dtype)
    out.index_add_(0, torch.as_tensor(dst), msg)
    return out


def union_layer(h, rel, g, w_n, w_loop, w_evolve, skip=None, prev_h=None):
    """rgcn/layers.py:222-279 (UnionRGCNLayer.forward/msg_func/apply_func), activation rrelu, eval mode.
    skip = (skip_connect_weight, skip_connect_bias) or None."""
    src, etype = torch.as_tensor(g["src"]), torch.as_tensor(g["etype"])
    msg = (h[src] + rel[etype]) @ w_n                                          # :257-276
    agg = scatter_sum(msg, g["dst"], g["num_nodes"]) * torch.as_tensor(g["norm"]).to(h.dtype).view(-1, 1)   # :278-279
    node = agg
    if w_loop is not None:
        has_in = torch.as_tensor(g["indeg"] > 0).view(-1, 1)
        loop = torch.where(has_in, h @ w_loop, h @ w_evolve)                   # :229-233
        node = node + loop
    if skip is not None and prev_h is not None and len(prev_h) != 0:
        sw = torch.sigmoid(prev_h @ skip[0] + skip[1])                          # :234-245
        node = sw * node + (1 - sw) * prev_h
    return rrelu(node)


def block_layer(h, g, weight, num_bases, out_feat):
    """rgcn/layers.py:167-179 (RGCNBlockLayer.msg_func/apply_func) through RGCNLayer.forward :48-91 with
    self_loop=False, no bias, activation rrelu (the static-graph configuration, src/rrgcn.py:104-105)."""
    si = h.shape[1] // num_bases
    so = out_feat // num_bases
    src, etype = torch.as_tensor(g["src"]), torch.as_tensor(g["etype"])
    w = weight[etype].view(-1, si, so)
    node = h[src].reshape(-1, 1, si)
    msg = torch.bmm(node, w).view(-1, out_feat)
    agg = scatter_sum(msg, g["dst"], g["num_nodes"]) * torch.as_tensor(g["norm"]).to(h.dtype).view(-1, 1)
    return rrelu(agg)


def regcn_static_emb(P, sg, num_ents, num_bases, layer_norm):
    """src/rrgcn.py:146-152: RGCNBlockLayer over the entity-word graph on cat(dynamic_emb, words_emb), entity rows,
    F.normalize when layer_norm.  (No dropout on this layer: RGCNLayer only drops the self-loop message, :52-53.)"""
    x = torch.cat((P["dynamic_emb"], P["words_emb"]), dim=0)
    out = block_layer(x, sg, P["statci_rgcn_layer.weight"], num_bases, x.shape[1])[:num_ents]
    return normalize_rows(out) if layer_norm else out


def static_angle_loss(static_emb, hist, layer_norm, angle, discount, weight):
    """src/rrgcn.py:225-247."""
    loss = static_emb.new_zeros(())
    for t, e in enumerate(hist):
        step = (angle * math.pi / 180) * ((t + 1) if discount == 1 else 1)
        if layer_norm:
            sim = (static_emb * normalize_rows(e)).sum(dim=1)
        else:
            sim = (static_emb * e).sum(dim=1) / (static_emb.norm(dim=1) * e.norm(dim=1))
        v = math.cos(step) - sim
        loss = loss + weight * v[v > 0].sum()
    return loss


def regcn_forward(p, graphs, num_rels, layer_norm=True, n_layers=2, self_loop=True, dtype=torch.float32,
                  trace=None, h_init=None):
    """src/rrgcn.py:142-180 (RecurrentRGCN.forward, use_static=False).  p: state dict (reference names).
    Returns (history_embs, h_0).  `trace` (dict) collects intermediates for per-kernel parity tests."""
    P = {k: v.to(dtype) for k, v in p.items() if v.is_floating_point()}
    h = normalize_rows(P["dynamic_emb"]) if layer_norm else P["dynamic_emb"]
    if h_init is not None:
        h = h_init                                  # use_static: the static embedding replaces it (src/rrgcn.py:152)
    emb_rel = P["emb_rel"]
    h0 = None
    hist = []
    for i, g in enumerate(graphs):
        rel_rowptr, rel_ents = r2e(g["triples"], num_rels)
        x_mean = rel_mean_pool(h, rel_rowptr, rel_ents, num_rels)
        x_in = torch.cat((emb_rel, x_mean), dim=1)
        h0 = gru_cell(x_in, emb_rel if i == 0 else h0, P["relation_cell_1.weight_ih"], P["relation_cell_1.weight_hh"],
                      P["relation_cell_1.bias_ih"], P["relation_cell_1.bias_hh"])
        h0 = normalize_rows(h0) if layer_norm else h0
        cur = h
        for l in range(n_layers):
            pre = f"rgcn.layers.{l}."
            cur = union_layer(cur, h0, g, P[pre + "weight_neighbor"],
                              P.get(pre + "loop_weight") if self_loop else None,
                              P.get(pre + "evolve_loop_weight") if self_loop else None)
            if trace is not None:
                trace[f"s{i}.l{l}.out"] = cur
        cur = normalize_rows(cur) if layer_norm else cur
        tw = torch.sigmoid(h @ P["time_gate_weight"] + P["time_gate_bias"])     # :177
        h = tw * cur + (1 - tw) * h                                              # :178
        hist.append(h)
        if trace is not None:
            trace[f"s{i}.x_mean"] = x_mean
            trace[f"s{i}.h0"] = h0
    return hist, h0


# =====================================================================================
# ConvTransE / ConvTransR (src/decoder.py:29-52, 78-100), eval-mode BatchNorm
# =====================================================================================
def _bn_eval(x, P, pre, eps=1e-5):
    shape = [1, -1] + [1] * (x.dim() - 2)
    return ((x - P[pre + "running_mean"].view(shape)) / torch.sqrt(P[pre + "running_var"].view(shape) + eps)
            * P[pre + "weight"].view(shape) + P[pre + "bias"].view(shape))


def _conv1d_same(x, w, b):
    """Conv1d(2->C, k, padding=k//2), stride 1: x (B,2,d), w (C,2,k)."""
    k = w.shape[2]
    xp = torch.nn.functional.pad(x, (k // 2, k // 2))
    cols = xp.unfold(2, k, 1)                                  # (B,2,d,k)
    return torch.einsum("bidk,cik->bcd", cols, w) + b.view(1, -1, 1)


def conv_tower(first, second, P, pre, always_bn2):
    x = torch.stack([first, second], dim=1)                    # (B,2,d)
    x = _bn_eval(x, P, pre + "bn0.")
    x = _conv1d_same(x, P[pre + "conv1.weight"], P[pre + "conv1.bias"])
    x = torch.relu(_bn_eval(x, P, pre + "bn1."))
    x = x.reshape(x.shape[0], -1)
    x = x @ P[pre + "fc.weight"].t() + P[pre + "fc.bias"]
    if always_bn2 or x.shape[0] > 1:
        x = _bn_eval(x, P, pre + "bn2.")
    return torch.relu(x)


def convtranse_scores(P, emb, rel, triples, pre="decoder_ob."):
    """src/decoder.py:78-100 (mode 'test'): `b` is registered but never added (:72)."""
    e_all = torch.tanh(emb)
    t = torch.as_tensor(triples)
    q = conv_tower(e_all[t[:, 0]], rel[t[:, 1]], P, pre, always_bn2=False)
    return q @ e_all.t()


def convtransr_scores(P, emb, rel, triples, pre="rdecoder."):
    """src/decoder.py:29-52."""
    e_all = torch.tanh(emb)
    t = torch.as_tensor(triples)
    q = conv_tower(e_all[t[:, 0]], e_all[t[:, 2]], P, pre, always_bn2=True)
    return q @ rel.t()


def add_inverse(test_triples, num_rels):
    """src/rrgcn.py:185-187."""
    t = np.asarray(test_triples, dtype=np.int64)
    inv = t[:, [2, 1, 0]].copy()
    inv[:, 1] += num_rels
    return np.concatenate((t, inv))


def regcn_predict(p, graphs, num_rels, test_triples, layer_norm=True, dtype=torch.float32, **kw):
    """src/rrgcn.py:183-194."""
    P = {k: v.to(dtype) for k, v in p.items() if v.is_floating_point()}
    all_triples = add_inverse(test_triples, num_rels)
    hist, h0 = regcn_forward(p, graphs, num_rels, layer_norm=layer_norm, dtype=dtype, **kw)
    emb = normalize_rows(hist[-1]) if layer_norm else hist[-1]
    return all_triples, convtranse_scores(P, emb, h0, all_triples), convtransr_scores(P, emb, h0, all_triples), hist, h0


# =====================================================================================
# Ranking (rgcn/utils.py:21-25, 51-75, 136-166)
# =====================================================================================
def stable_rank0(score, target):
    """0-based position of the target after a *stable* descending sort: #{s_j > s_t} + #{j < t, s_j == s_t}.
    Equals the reference's torch.sort position whenever the target score is unique in its row."""
    score = np.asarray(score)
    target = np.asarray(target, dtype=np.int64)
    st = score[np.arange(score.shape[0]), target][:, None]
    idx = np.arange(score.shape[1])[None, :]
    return ((score > st) | ((score == st) & (idx < target[:, None]))).sum(axis=1).astype(np.int64)


def filter_scores(triples, score, all_ans, rel_predict=0):
    """rgcn/utils.py:51-75 on a copy."""
    score = np.array(score, copy=True)
    for i, (h, r, t) in enumerate(np.asarray(triples)):
        if rel_predict:
            ans = set(all_ans[int(h)][int(t)])
            ans.discard(int(r))
        else:
            ans = set(all_ans[int(h)][int(r)])
            ans.discard(int(t))
        if ans:
            score[i, sorted(ans)] = FILTER_SCORE
    return score


def total_rank(triples, score, all_ans, rel_predict=0):
    """rgcn/utils.py:136-166: returns (filter_mrr, mrr, rank, filter_rank), ranks 1-based int64."""
    triples = np.asarray(triples)
    target = triples[:, {0: 2, 1: 1, 2: 0}[rel_predict]]
    rank = stable_rank0(score, target) + 1
    fscore = filter_scores(triples, score, all_ans, rel_predict) if all_ans is not None else score
    frank = stable_rank0(fscore, target) + 1
    mrr = float(np.mean(np.float32(1.0) / rank.astype(np.float32)))
    fmrr = float(np.mean(np.float32(1.0) / frank.astype(np.float32)))
    return fmrr, mrr, rank, frank


# =====================================================================================
# Poincare / Lorentz primitives (hyperbolic_ops.py)
# =====================================================================================
def _norm(x):
    return x.norm(p=2, dim=-1, keepdim=True)


def clamp_norm(x, max_norm, eps=EPS):
    """hyperbolic_ops.py:38-53."""
    n = _norm(x).clamp(min=eps)
    return x * (torch.clamp(n, max=max_norm - eps) / n)


def project(x, c, eps=EPS):
    """hyperbolic_ops.py:56-74."""
    return clamp_norm(x, 1.0 / math.sqrt(c) - eps, eps)


def exp0(v, c, eps=EPS):
    """hyperbolic_ops.py:77-95."""
    sc = math.sqrt(c)
    n = _norm(v).clamp(min=eps)
    return project(torch.tanh(sc * n) * (v / n) / sc, c, eps)


def log0(x, c, eps=EPS):
    """hyperbolic_ops.py:98-116."""
    sc = math.sqrt(c)
    n = _norm(x).clamp(min=eps)
    return torch.atanh((sc * n).clamp(max=1.0 - eps)) * x / (sc * n)


def mobius_add(x, y, c, eps=EPS):
    """hyperbolic_ops.py:119-143."""
    x_sq = (x * x).sum(-1, keepdim=True)
    y_sq = (y * y).sum(-1, keepdim=True)
    xy = (x * y).sum(-1, keepdim=True)
    num = (1 + 2 * c * xy + c * y_sq) * x + (1 - c * x_sq) * y
    den = 1 + 2 * c * xy + c * c * x_sq * y_sq
    return project(num / (den + eps), c, eps)


def get_radius(x, eps=EPS):
    """hyperbolic_ops.py:194-206."""
    return x.norm(p=2, dim=-1).clamp(min=eps)


def apply_radius(x, radius, c, eps=EPS):
    """hyperbolic_ops.py:209-233."""
    r = radius.unsqueeze(-1) if radius.dim() == x.dim() - 1 else radius
    r = r.clamp(min=eps, max=1.0 / math.sqrt(c) - eps)
    return x / _norm(x).clamp(min=eps) * r


def to_lorentz(x, c, eps=EPS):
    """hyperbolic_ops.py:477-499."""
    sc = math.sqrt(c)
    nsq = (x ** 2).sum(-1, keepdim=True)
    den = (1.0 - c * nsq).clamp(min=eps)
    return torch.cat([(1.0 + c * nsq) / (sc * den), 2.0 * x / den], dim=-1)


def to_poincare(y, c, eps=EPS):
    """hyperbolic_ops.py:502-518."""
    return y[..., 1:] / (1.0 + y[..., :1] * math.sqrt(c)).clamp(min=eps)


def lorentz_centroid(emb, w, c, eps=EPS):
    """hyperbolic_ops.py:563-581."""
    w = w / (w.sum() + eps)
    cen = (w.unsqueeze(-1) * emb).sum(0)
    ip = -(cen[:1] * cen[:1]).sum(-1, keepdim=True) + (cen[1:] * cen[1:]).sum(-1, keepdim=True)
    return cen / torch.sqrt(torch.clamp(-ip * c, min=eps))


def static_radius(radius_static, c, rmin, rmax):
    """hyperbolic_model.py:715-720."""
    r = radius_static.clamp(min=rmin, max=rmax)
    return r.clamp(max=1.0 / math.sqrt(c) - 1e-6)


def radius_evolution(x, rs, w, b, c, beta, eps_r):
    """hyperbolic_ops.py:395-435 (TemporalRadiusEvolution.forward)."""
    delta = (log0(x, c) @ w.t() + b).squeeze(-1).clamp(min=-eps_r, max=eps_r)
    base = beta * rs + (1.0 - beta) * get_radius(x)
    return apply_radius(x, base + delta, c)


# =====================================================================================
# Hyperbolic encoders (hyperbolic_layers.py) and recurrent model (hyperbolic_model.py)
# =====================================================================================
def hyp_union_layer(h_hyper, rel, g, w_n, w_loop, w_evolve, c, gamma):
    """hyperbolic_layers.py:222-323 (HyperbolicUnionRGCNLayer, no skip connection, rrelu, eval)."""
    ht = log0(h_hyper, c)
    radius = get_radius(h_hyper).unsqueeze(-1)
    src, dst, etype = (torch.as_tensor(g[k]) for k in ("src", "dst", "etype"))
    msg = (ht[src] + rel[etype]) @ w_n                                           # :225-231
    msg = msg * torch.exp(-gamma * (radius[src] - radius[dst]).abs())           # :232-234
    agg = scatter_sum(msg, g["dst"], g["num_nodes"]) * torch.as_tensor(g["norm"]).to(ht.dtype).view(-1, 1)
    h_new = agg.clamp(-10.0, 10.0)                                               # :296
    if w_loop is not None:
        has_in = torch.as_tensor(g["indeg"] > 0).view(-1, 1)
        h_new = h_new + torch.where(has_in, ht @ w_loop, ht @ w_evolve)          # :273-282, :306-307
    h_new = rrelu(h_new.clamp(-10.0, 10.0))                                      # :310-314
    return exp0(h_new, c)                                                        # :321


def hyp_rgcn_layer(h_hyper, g, weight, num_bases, c, gamma, w_loop=None, skip=None, prev_h=None, act=True):
    """hyperbolic_layers.py:87-161 (HyperbolicRGCNLayer.msg_func / apply_func / forward, eval): block-diagonal relation
    transform of log_0(h[src]) weighted by exp(-gamma |r_src - r_dst|), sum, degree norm, + self loop, skip gate on
    log_0(prev_h), activation, exp_0.  skip = (skip_weight, skip_bias)."""
    ht = log0(h_hyper, c)
    radius = get_radius(h_hyper).unsqueeze(-1)
    out_feat = ht.shape[1]
    si, so = ht.shape[1] // num_bases, out_feat // num_bases
    src, dst, etype = (torch.as_tensor(g[k]) for k in ("src", "dst", "etype"))
    w = weight[etype].view(-1, si, so)                                            # :90-91
    msg = torch.bmm(ht[src].reshape(-1, 1, si), w).view(-1, out_feat)             # :94-98
    msg = msg * torch.exp(-gamma * (radius[src] - radius[dst]).abs())             # :99-101
    h_new = scatter_sum(msg, g["dst"], g["num_nodes"]) * torch.as_tensor(g["norm"]).to(ht.dtype).view(-1, 1)   # :107-109
    if w_loop is not None:
        h_new = h_new + ht @ w_loop                                               # :138-140
    if skip is not None and prev_h is not None:
        pt = log0(prev_h, c)
        gate = torch.sigmoid(pt @ skip[0] + skip[1])                              # :143-146
        h_new = gate * h_new + (1 - gate) * pt
    if act:
        h_new = rrelu(h_new)                                                      # :149-150
    return exp0(h_new, c)                                                         # :157


def lorentz_layer(h_hyper, rel, g, weight, w_loop, w_evolve, c, num_bases, skip=None, prev_h=None):
    """hyperbolic_layers.py:589-694 (LorentzRGCNLayer, rrelu, eval).  skip = (skip_weight, skip_bias) with prev_h = the
    previous layer's INPUT (hyperbolic_layers.py:657-662, 675-678, cell :737-740)."""
    n, d = h_hyper.shape
    nb = num_bases
    sb = d // nb
    ht = log0(h_hyper, c)
    src, etype = torch.as_tensor(g["src"]), torch.as_tensor(g["etype"])
    w = weight[etype].view(-1, sb, sb)
    m = torch.bmm(ht[src].reshape(-1, 1, sb), w).view(-1, d)                      # :593-599
    if rel is not None:
        m = m + rel[etype]                                                       # :602-606
    m_l = to_lorentz(exp0(m, c), c)                                              # :609-610
    norm = torch.as_tensor(g["norm"]).to(ht.dtype)
    out = torch.zeros(n, d + 1, dtype=ht.dtype)
    rowptr, eperm, _, _ = csr_by_dst(g)
    for v in range(n):
        b, e = int(rowptr[v]), int(rowptr[v + 1])
        if e == b:
            continue
        k = e - b
        nv = norm[v].repeat(k)
        wts = nv / (nv.sum() + 1e-6)                                             # :620
        out[v] = lorentz_centroid(m_l[torch.as_tensor(eperm[b:e])], wts, c)       # :622-624
    h_new = log0(to_poincare(out, c), c).clamp(-10.0, 10.0)                      # :670-672
    if w_loop is not None:
        has_in = torch.as_tensor(g["indeg"] > 0).view(-1, 1)
        h_new = h_new + torch.where(has_in, ht @ w_loop, ht @ w_evolve)          # :649-655, :681
    if skip is not None and prev_h is not None:
        pt = log0(prev_h, c)
        gate = torch.sigmoid(pt @ skip[0] + skip[1])
        h_new = gate * h_new + (1 - gate) * pt                                   # :678
    h_new = rrelu(h_new.clamp(-10.0, 10.0))                                      # :683-687
    return exp0(h_new, c)                                                        # :694


def hyp_forward(p, graphs, num_rels, c=0.01, encoder="hyperbolic_uvrgcn", layer_norm=False, n_layers=2,
                self_loop=True, gamma=1.0, num_bases=100, rmin=0.5, rmax=3.0, beta=1.0, eps_r=0.1, residual=True,
                dtype=torch.float32, trace=None, static_init=None):
    """hyperbolic_model.py:722-890 (HyperbolicRecurrentRGCN.forward; no geoopt, no EST).  static_init: the (already
    normalised) static embedding that replaces the initial table with --add-static-graph (:762-771)."""
    P = {k: v.to(dtype) for k, v in p.items() if v.is_floating_point()}
    init = normalize_rows(P["dynamic_emb"]) if layer_norm else P["dynamic_emb"]
    if static_init is not None:
        init = static_init
    h = exp0(init, c)                                                            # :779-780
    rs = static_radius(P["radius_static"], c, rmin, rmax)
    h = apply_radius(h, rs, c)                                                   # :782
    emb_rel = P["emb_rel"]
    h0 = None
    hist = []
    for i, g in enumerate(graphs):
        ht = log0(h, c)                                                          # :802
        rel_rowptr, rel_ents = r2e(g["triples"], num_rels)
        x_mean = rel_mean_pool(ht, rel_rowptr, rel_ents, num_rels)               # :803-812
        x_in = torch.cat((emb_rel, x_mean), dim=1)
        h0 = gru_cell(x_in, emb_rel if i == 0 else h0, P["relation_gru.weight_ih"], P["relation_gru.weight_hh"],
                      P["relation_gru.bias_ih"], P["relation_gru.bias_hh"])      # :815-824
        h0 = normalize_rows(h0) if layer_norm else h0
        cur = h
        prev_in = None
        for l in range(n_layers):
            pre = f"rgcn.layers.{l}."
            wl = P.get(pre + "loop_weight") if self_loop else None
            we = P.get(pre + "evolve_loop_weight") if self_loop else None
            if encoder == "hyperbolic_uvrgcn":
                cur = hyp_union_layer(cur, h0, g, P[pre + "weight_neighbor"], wl, we, c, gamma)
            elif encoder == "lgcn":
                skip = (P[pre + "skip_weight"], P[pre + "skip_bias"]) if pre + "skip_weight" in P else None
                cur, prev_in = lorentz_layer(cur, h0, g, P[pre + "weight"], wl, we, c, num_bases, skip, prev_in), cur
            else:
                raise NotImplementedError(encoder)
            if trace is not None:
                trace[f"s{i}.l{l}.out"] = cur
        cur = project(cur, c)                                                    # :829
        if layer_norm:
            cur = exp0(normalize_rows(log0(cur, c)), c)                          # :832-835
        ct = log0(cur, c).clamp(-10.0, 10.0)                                     # :841-846
        pt = log0(h, c).clamp(-10.0, 10.0)
        tw = torch.sigmoid(pt @ P["time_gate_weight"] + P["time_gate_bias"])     # :848
        h = project(exp0(tw * ct + (1 - tw) * pt, c), c)                         # :849-860
        if residual:
            h = radius_evolution(h, rs, P["temporal_radius_evolution.radius_mlp.weight"],
                                 P["temporal_radius_evolution.radius_mlp.bias"], c, beta, eps_r)   # :866-867
        else:
            h = apply_radius(h, rs, c)
        hist.append(h)
        if trace is not None:
            trace[f"s{i}.h0"] = h0
    return hist, h0


# =====================================================================================
# Hyperbolic decoders (hyperbolic_decoder.py)
# =====================================================================================
def hyp_dist_scores(query, cand, bias, c, scale, margin, chunk=4096, query_curvature=None):
    """hyperbolic_decoder.py:89-179.  Proxy-distance branch (use_hyperbolic_distance=False, :164-167):
    scale * (margin - |(-q) (+)_c e|^2) + bias; with `query_curvature` (B,) the true-distance branch with a
    per-query curvature (:145-163): scale * (margin - 2/sqrt(c_q) atanh(sqrt(c_q) |(-q) (+)_{c_q} e|)) + bias.
    Evaluated pair by pair in d dimensions (chunked over queries only)."""
    out = torch.empty(query.shape[0], cand.shape[0], dtype=query.dtype)
    step = max(1, min(64, chunk // max(1, cand.shape[0] // 64)))
    for b0 in range(0, query.shape[0], step):
        q = query[b0:b0 + step]
        qe = q.unsqueeze(1).expand(-1, cand.shape[0], -1)
        ce = cand.unsqueeze(0).expand(q.shape[0], -1, -1)
        if query_curvature is None:
            diff = mobius_add(-qe, ce, c)
            blk = scale * (margin - (diff ** 2).sum(-1))
        else:
            c_eff = query_curvature[b0:b0 + step].reshape(-1, 1, 1).to(query.dtype)
            sqrt_c = torch.sqrt(c_eff + EPS)
            x_sq = (qe * qe).sum(-1, keepdim=True)
            y_sq = (ce * ce).sum(-1, keepdim=True)
            xy = (qe * ce).sum(-1, keepdim=True)
            num = (1 - 2 * c_eff * xy + c_eff * y_sq) * (-qe) + (1 - c_eff * x_sq) * ce
            denom = 1 - 2 * c_eff * xy + (c_eff ** 2) * x_sq * y_sq
            diff = num / (denom + EPS)
            dn = torch.norm(diff, p=2, dim=-1, keepdim=True).clamp(min=EPS)
            dn = torch.min(dn, 1.0 / (sqrt_c + EPS) - EPS)
            dist = (2.0 / (sqrt_c + EPS)) * torch.atanh((sqrt_c * dn).clamp(max=1.0 - EPS))
            blk = scale * (margin - dist.squeeze(-1))
        if bias is not None:
            blk = blk + bias.unsqueeze(0)
        out[b0:b0 + step] = blk
    return out


def relation_curvature(P, pre, r_idx, c, cmax=None):
    """hyperbolic_decoder.py:66-86,1020-1026: c_q = max(1e-5, min(softplus(raw[r mod R]), 0.999 c, cmax)); None
    when the decoder has no rel_curvature_raw."""
    raw = P.get(pre + "rel_curvature_raw")
    if raw is None:
        return None
    rel_c = _softplus(raw[torch.remainder(r_idx, raw.shape[0])])
    upper = raw.new_tensor(0.999 * float(c))
    if cmax is not None:
        upper = torch.min(upper, raw.new_tensor(float(cmax)))
    return torch.max(torch.min(rel_c, upper), raw.new_tensor(1e-5))


def _flagged_scores(P, pre, q, emb, t, c, scale):
    """Shared tail of RotH / MuRP (:1087-1099, :767-779): candidate bias, per-query curvature, subject bias."""
    eb = P.get(pre + "entity_bias")
    rel_c = relation_curvature(P, pre, t[:, 1], c, cmax=c)
    s = hyp_dist_scores(q, emb, eb, c, scale, P[pre + "score_margin"], query_curvature=rel_c)
    if eb is not None:
        s = s + eb[t[:, 0]].unsqueeze(1)
    return s


def givens(x, ang):
    """hyperbolic_decoder.py:1033-1051."""
    x1, x2 = x[:, 0::2], x[:, 1::2]
    ca, sa = torch.cos(ang), torch.sin(ang)
    return torch.stack([ca * x1 - sa * x2, sa * x1 + ca * x2], dim=2).reshape(x.shape)


def _softplus(x):
    return torch.nn.functional.softplus(x)


def roth_scores(P, emb, rel, triples, c, pre="decoder_ob."):
    """hyperbolic_decoder.py:1053-1099 (HyperbolicRotH.forward, eval; optional entity bias / relation curvature)."""
    t = torch.as_tensor(triples)
    r_idx = t[:, 1]
    s_tan = log0(project(emb[t[:, 0]], c), c)
    h1 = torch.relu(s_tan @ P[pre + "reshape_fc1.weight"].t() + P[pre + "reshape_fc1.bias"])
    s_tan = s_tan + (h1 @ P[pre + "reshape_fc2.weight"].t() + P[pre + "reshape_fc2.bias"])      # :1028-1030
    ang = rel[r_idx] @ P[pre + "rot_proj.weight"].t() + P[pre + "rot_proj.bias"]
    rot_s = project(exp0(givens(s_tan, ang), c), c)
    v_r = rel[r_idx] @ P[pre + "trans_proj.weight"].t() + P[pre + "trans_proj.bias"]
    t_r = project(exp0(v_r, c), c)
    q = mobius_add(rot_s, t_r, c)
    scale = _softplus(P[pre + "score_scale_raw"]) + 1e-6
    return _flagged_scores(P, pre, q, emb, t, c, scale), q


def givens_reflection(x, ang):
    """hyperbolic_decoder.py:1391-1401: pairs -> (cos a x1 + sin a x2, sin a x1 - cos a x2)."""
    x1, x2 = x[:, 0::2], x[:, 1::2]
    ca, sa = torch.cos(ang), torch.sin(ang)
    return torch.stack([ca * x1 + sa * x2, sa * x1 - ca * x2], dim=2).reshape(x.shape)


def atth_scores(P, emb, rel, triples, c, pre="decoder_ob."):
    """hyperbolic_decoder.py:1403-1480 (HyperbolicAttH.forward, eval)."""
    t = torch.as_tensor(triples)
    r = rel[t[:, 1]]
    s_tan = log0(project(emb[t[:, 0]], c), c)
    lin = lambda n, x: x @ P[pre + n + ".weight"].t() + P[pre + n + ".bias"]
    rot_s = givens(s_tan, lin("rot_proj", r))
    ref_s = givens_reflection(s_tan, lin("ref_proj", r))
    a = torch.sigmoid((lin("attn_proj", r) * torch.cat([s_tan, r], dim=-1)).sum(dim=-1, keepdim=True))
    mixed = project(exp0(a * rot_s + (1.0 - a) * ref_s, c), c)
    t_r = project(exp0(lin("trans_proj", r), c), c)
    q = mobius_add(mixed, t_r, c)
    scale = _softplus(P[pre + "score_scale_raw"]) + 1e-6
    return _flagged_scores(P, pre, q, emb, t, c, scale), q


def atthrel_scores(P, emb, rel, triples, c, pre="rdecoder."):
    """hyperbolic_decoder.py:1593-1640 (HyperbolicAttHRel.forward)."""
    t = torch.as_tensor(triples)
    o_emb = emb[t[:, 2]]
    s_tan, o_tan = log0(emb[t[:, 0]], c), log0(o_emb, c)
    n = s_tan.shape[0]
    rot_s = givens(s_tan, P[pre + "global_rot"].unsqueeze(0).expand(n, -1))
    ref_s = givens_reflection(s_tan, P[pre + "global_ref"].unsqueeze(0).expand(n, -1))
    a = torch.sigmoid(torch.cat([s_tan, o_tan], dim=-1) @ P[pre + "attn_weight"]).unsqueeze(1)
    q = mobius_add(-exp0(a * rot_s + (1.0 - a) * ref_s, c), o_emb, c)
    scale = _softplus(P[pre + "score_scale_raw"]) + 1e-6
    return hyp_dist_scores(q, exp0(rel, c), P[pre + "rel_bias"], c, scale, P[pre + "score_margin"]), q


def murp_scores(P, emb, rel, triples, c, pre="decoder_ob."):
    """hyperbolic_decoder.py:733-779 (HyperbolicMuRP.forward)."""
    t = torch.as_tensor(triples)
    r_idx = t[:, 1]
    s_emb = project(emb[t[:, 0]], c)
    rot = rel[r_idx] @ P[pre + "rot_proj.weight"].t() + P[pre + "rot_proj.bias"]
    rot_s = project(exp0(rot * log0(s_emb, c), c), c)
    v_r = rel[r_idx] @ P[pre + "trans_proj.weight"].t() + P[pre + "trans_proj.bias"]
    t_r = project(exp0(v_r, c), c)
    q = mobius_add(rot_s, t_r, c)
    scale = _softplus(P[pre + "score_scale_raw"]) + 1e-6
    return _flagged_scores(P, pre, q, emb, t, c, scale), q


def rothrel_scores(P, emb, rel, triples, c, pre="rdecoder."):
    """hyperbolic_decoder.py:1230-1262 (HyperbolicRotHRel.forward)."""
    t = torch.as_tensor(triples)
    s_tan = log0(emb[t[:, 0]], c)
    h1 = torch.relu(s_tan @ P[pre + "reshape_fc1.weight"].t() + P[pre + "reshape_fc1.bias"])
    s_tan = s_tan + (h1 @ P[pre + "reshape_fc2.weight"].t() + P[pre + "reshape_fc2.bias"])
    ang = P[pre + "global_rot"].unsqueeze(0).expand(s_tan.shape[0], -1)
    rot_s = exp0(givens(s_tan, ang), c)
    q = mobius_add(-rot_s, emb[t[:, 2]], c)
    rel_hyp = exp0(rel, c)
    scale = _softplus(P[pre + "score_scale_raw"]) + 1e-6
    return hyp_dist_scores(q, rel_hyp, P[pre + "rel_bias"], c, scale, P[pre + "score_margin"]), q


def murprel_scores(P, emb, rel, triples, c, pre="rdecoder."):
    """hyperbolic_decoder.py:856-882 (HyperbolicMuRPRel.forward): no scale, margin 0."""
    t = torch.as_tensor(triples)
    q_tan = log0(emb[t[:, 0]], c) @ P[pre + "W_s"] + log0(emb[t[:, 2]], c) @ P[pre + "W_o"]
    q = exp0(q_tan, c)
    one = torch.ones((), dtype=emb.dtype)
    return hyp_dist_scores(q, exp0(rel, c), P[pre + "rel_bias"], c, one, 0.0 * one), q


def hyp_convtranse_scores(P, emb, rel, triples, c, pre="decoder_ob."):
    """hyperbolic_decoder.py:360-413."""
    et = log0(emb, c)
    et = 0.9 * torch.tanh(et) + 0.1 * et
    t = torch.as_tensor(triples)
    q = conv_tower(et[t[:, 0]], rel[t[:, 1]], P, pre, always_bn2=False)
    return q @ et.t() + P[pre + "b"]


def hyp_convtransr_scores(P, emb, rel, triples, c, pre="rdecoder."):
    """hyperbolic_decoder.py:464-510."""
    et = log0(emb, c)
    et = 0.9 * torch.tanh(et) + 0.1 * et
    t = torch.as_tensor(triples)
    q = conv_tower(et[t[:, 0]], et[t[:, 2]], P, pre, always_bn2=True)
    return q @ rel.t() + P[pre + "b"]


def hyp_predict(p, graphs, num_rels, test_triples, c=0.01, decoder="roth", layer_norm=False, dtype=torch.float32,
                **kw):
    """hyperbolic_model.py:892-939."""
    P = {k: v.to(dtype) for k, v in p.items() if v.is_floating_point()}
    all_triples = add_inverse(test_triples, num_rels)
    hist, h0 = hyp_forward(p, graphs, num_rels, c=c, layer_norm=layer_norm, dtype=dtype, **kw)
    emb = hist[-1]
    if layer_norm:
        emb = exp0(normalize_rows(log0(emb, c)), c)
    if decoder == "roth":
        score = roth_scores(P, emb, h0, all_triples, c)[0]
        score_rel = rothrel_scores(P, emb, h0, all_triples, c)[0]
    elif decoder == "hyperbolic_convtranse":
        score = hyp_convtranse_scores(P, emb, h0, all_triples, c)
        score_rel = hyp_convtransr_scores(P, emb, h0, all_triples, c)
    elif decoder == "murp":
        score = murp_scores(P, emb, h0, all_triples, c)[0]
        score_rel = murprel_scores(P, emb, h0, all_triples, c)[0]
    elif decoder == "atth":
        score = atth_scores(P, emb, h0, all_triples, c)[0]
        score_rel = atthrel_scores(P, emb, h0, all_triples, c)[0]
    else:
        raise NotImplementedError(decoder)
    return all_triples, score, score_rel, hist, h0


# =====================================================================================
# Loss heads (eval-mode forward): src/rrgcn.py:197-223, hyperbolic_model.py:941-1036,1066-1073
# =====================================================================================
def cross_entropy(score, target):
    """nn.CrossEntropyLoss (mean): logsumexp(row) - row[target]; the hyperbolic decoders' streaming `loss`
    (hyperbolic_decoder.py:182-307) is the same quantity computed chunk by chunk."""
    t = torch.as_tensor(np.asarray(target), dtype=torch.long)
    return (torch.logsumexp(score, dim=1) - score[torch.arange(score.shape[0]), t]).mean()


def regcn_loss(p, graphs, num_rels, triples, layer_norm=True, dtype=torch.float32, **kw):
    """RecurrentRGCN.get_loss without the static-graph term: (loss_ent, loss_rel, 0)."""
    all_t, score, score_rel, _, _ = regcn_predict(p, graphs, num_rels, triples, layer_norm=layer_norm, dtype=dtype, **kw)
    return (float(cross_entropy(score, all_t[:, 2])), float(cross_entropy(score_rel, all_t[:, 1])), 0.0)


def hyp_loss(p, graphs, num_rels, triples, c=0.01, decoder="roth", layer_norm=False, radius_lambda=0.02, rmin=0.5,
             rmax=3.0, dtype=torch.float32, **kw):
    """HyperbolicRecurrentRGCN.get_loss: (loss_ent, loss_rel, 0, loss_radius); the per-query entity bias cancels in CE
    (hyperbolic_decoder.py:205-206) and radius supervision is an MSE over the entities of the batch (:1066-1073)."""
    all_t, score, score_rel, _, _ = hyp_predict(p, graphs, num_rels, triples, c=c, decoder=decoder,
                                                layer_norm=layer_norm, dtype=dtype, **kw)
    ids = torch.unique(torch.as_tensor(all_t[:, [0, 2]].reshape(-1)))
    rs = static_radius(p["radius_static"].to(dtype), c, rmin, rmax)[ids]
    rt = p["radius_target"].to(dtype)[ids]
    return (float(cross_entropy(score, all_t[:, 2])), float(cross_entropy(score_rel, all_t[:, 1])), 0.0,
            float(radius_lambda * torch.mean((rs - rt) ** 2)))


def construct_snap(all_triples, num_rels, score, topk, rel_mode=0):
    """rgcn/utils.py:367-405 (construct_snap / construct_snap_r): top-k of every row in descending order (ties by
    ascending id: the order of a stable sort; torch.sort leaves it unspecified), turned into predicted triples."""
    score = np.asarray(score)
    top = np.argsort(-score, axis=1, kind="stable")[:, :topk]
    out = []
    for q, (h, r, t) in enumerate(np.asarray(all_triples)):
        for idx in top[q]:
            if not rel_mode:
                out.append([h, r, idx] if r < num_rels else [idx, r - num_rels, h])
            else:
                out.append([h, idx, t] if idx < num_rels else [t, idx - num_rels, h])
    return np.asarray(out, dtype=np.int64).reshape(-1, 3)


# =====================================================================================
# Training step (src/rrgcn.py:197-223 in train() mode + src/main.py:235-246), dropout 0.
# Gradients come from torch autograd over the restated forward (CPU); BatchNorm uses batch statistics
# and updates its running ones like nn.BatchNorm1d (momentum 0.1, unbiased variance).
# =====================================================================================
def _bn_train(x, P, pre, stats, eps=1e-5, momentum=0.1):
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    n = x.numel() // x.shape[1]
    rm = stats.get(pre + "running_mean", P[pre + "running_mean"])
    rv = stats.get(pre + "running_var", P[pre + "running_var"])
    stats[pre + "running_mean"] = ((1 - momentum) * rm + momentum * mean).detach()
    stats[pre + "running_var"] = ((1 - momentum) * rv + momentum * var * n / max(n - 1, 1)).detach()
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * P[pre + "weight"].view(shape) + P[pre + "bias"].view(shape)


def conv_tower_train(first, second, P, pre, stats):
    """src/decoder.py:35-50 / 83-95 in train() mode with every dropout probability 0."""
    x = torch.stack([first, second], dim=1)
    x = _bn_train(x, P, pre + "bn0.", stats)
    x = _conv1d_same(x, P[pre + "conv1.weight"], P[pre + "conv1.bias"])
    x = torch.relu(_bn_train(x, P, pre + "bn1.", stats))
    x = x.reshape(x.shape[0], -1)
    x = x @ P[pre + "fc.weight"].t() + P[pre + "fc.bias"]
    x = _bn_train(x, P, pre + "bn2.", stats)
    return torch.relu(x)


def regcn_train_losses(P, graphs, num_rels, triples, layer_norm, stats, n_layers=2, static=None):
    """(loss_ent, loss_rel[, loss_static]) with the autograd tape on; P holds leaf tensors (parameters) and buffers.
    static = dict(graph, num_ents, num_bases, angle, discount, weight) switches the static-graph constraint on."""
    all_t = torch.as_tensor(add_inverse(triples, num_rels))
    s_emb = None
    if static is not None:
        s_emb = regcn_static_emb(P, static["graph"], static["num_ents"], static["num_bases"], layer_norm)
    hist, h0 = regcn_forward(P, graphs, num_rels, layer_norm=layer_norm, n_layers=n_layers, dtype=P["emb_rel"].dtype,
                             h_init=s_emb)
    emb = normalize_rows(hist[-1]) if layer_norm else hist[-1]
    e_all = torch.tanh(emb)
    q = conv_tower_train(e_all[all_t[:, 0]], h0[all_t[:, 1]], P, "decoder_ob.", stats)
    loss_e = cross_entropy(q @ e_all.t(), all_t[:, 2])
    q = conv_tower_train(e_all[all_t[:, 0]], e_all[all_t[:, 2]], P, "rdecoder.", stats)
    loss_r = cross_entropy(q @ h0.t(), all_t[:, 1])
    if static is not None:
        return loss_e, loss_r, static_angle_loss(s_emb, hist, layer_norm, static["angle"], static["discount"],
                                                 static["weight"])
    return loss_e, loss_r


def regcn_train_steps(sd, graphs, num_rels, triples, layer_norm=True, steps=1, task_weight=0.7, grad_norm=1.0,
                      lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, dtype=torch.float32, static=None):
    """`steps` optimisation steps (loss.backward, clip_grad_norm_, torch.optim.Adam with L2 weight decay) on the same
    batch.  Returns a list of per-step dicts {losses, grad_norm, grads{name}, params{name}} and the final buffers."""
    P = {}
    for k, v in sd.items():
        if not v.is_floating_point() or k == "rgcn.rel_emb":
            continue
        t = v.detach().clone().to(dtype)
        if "running_" not in k:
            t.requires_grad_(True)
        P[k] = t
    m = {k: torch.zeros_like(v) for k, v in P.items() if v.requires_grad}
    vv = {k: torch.zeros_like(v) for k, v in P.items() if v.requires_grad}
    log = []
    for step in range(1, steps + 1):
        stats = {}
        ls = regcn_train_losses(P, graphs, num_rels, triples, layer_norm, stats, static=static)
        loss_e, loss_r = ls[0], ls[1]
        loss = task_weight * loss_e + (1 - task_weight) * loss_r
        if static is not None:
            loss = loss + ls[2]
        names = [k for k, v in P.items() if v.requires_grad]
        gs = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
        grads = {k: g for k, g in zip(names, gs) if g is not None}
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
        coef = min(1.0, grad_norm / (float(total) + 1e-6))
        with torch.no_grad():
            for k, g in grads.items():
                p = P[k]
                gg = g * coef + weight_decay * p
                m[k] = betas[0] * m[k] + (1 - betas[0]) * gg
                vv[k] = betas[1] * vv[k] + (1 - betas[1]) * gg * gg
                denom = vv[k].sqrt() / math.sqrt(1 - betas[1] ** step) + eps
                p -= (lr / (1 - betas[0] ** step)) * m[k] / denom
            for k, s in stats.items():
                P[k] = s
        log.append({"losses": tuple(float(x.detach()) for x in ls), "grad_norm": float(total),
                    "grads": {k: g.detach().clone() for k, g in grads.items()},
                    "params": {k: P[k].detach().clone() for k in grads}})
    return log, {k: v.detach() for k, v in P.items() if "running_" in k}


# =====================================================================================
# Hyperbolic training step (hyperbolic_model.py:941-1088 in train() mode + hyperbolic_main.py:585-628), dropout 0,
# hyperbolic_uvrgcn encoder + hyperbolic_convtranse decoder.
# =====================================================================================
def hyp_static_emb(P, sg, num_ents, num_bases, layer_norm):
    """hyperbolic_model.py:762-770: the block layer is called `static_rgcn_layer` here."""
    x = torch.cat((P["dynamic_emb"], P["words_emb"]), dim=0)
    out = block_layer(x, sg, P["static_rgcn_layer.weight"], num_bases, x.shape[1])[:num_ents]
    return normalize_rows(out) if layer_norm else out


def hyp_train_losses(P, graphs, num_rels, triples, c, layer_norm, gamma, stats, rmin=0.5, rmax=3.0, beta=1.0, eps_r=0.1,
                     radius_lambda=0.02, decoder="hyperbolic_convtranse", encoder="hyperbolic_uvrgcn", num_bases=100,
                     static=None):
    all_t = torch.as_tensor(add_inverse(triples, num_rels))
    s_emb = None
    if static is not None:
        s_emb = hyp_static_emb(P, static["graph"], static["num_ents"], static["num_bases"], layer_norm)
    hist, h0 = hyp_forward(P, graphs, num_rels, c=c, encoder=encoder, layer_norm=layer_norm, gamma=gamma,
                           num_bases=num_bases, rmin=rmin, rmax=rmax, beta=beta, eps_r=eps_r, dtype=P["emb_rel"].dtype,
                           static_init=s_emb)
    emb = hist[-1]
    if layer_norm:
        emb = exp0(normalize_rows(log0(emb, c)), c)
    if decoder in ("murp", "roth", "atth"):
        # Hyperbolic{MuRP,RotH,AttH}.loss / ...Rel.loss (hyperbolic_decoder.py:781-817, 897-928, 1101-1138, 1264-1280,
        # 1482-1512, 1642-1700):
        # CE over the forward scores; the per-query subject bias of forward() is not part of loss() (and cancels in CE)
        ent_fn, rel_fn = {"murp": (murp_scores, murprel_scores), "roth": (roth_scores, rothrel_scores),
                          "atth": (atth_scores, atthrel_scores)}[decoder]
        sc = ent_fn(P, emb, h0, all_t.numpy(), c)[0]
        if "decoder_ob.entity_bias" in P:
            sc = sc - P["decoder_ob.entity_bias"][all_t[:, 0]].unsqueeze(1)
        loss_e = cross_entropy(sc, all_t[:, 2])
        loss_r = cross_entropy(rel_fn(P, emb, h0, all_t.numpy(), c)[0], all_t[:, 1])
    else:
        et = log0(emb, c)
        et = 0.9 * torch.tanh(et) + 0.1 * et
        q = conv_tower_train(et[all_t[:, 0]], h0[all_t[:, 1]], P, "decoder_ob.", stats)
        loss_e = cross_entropy(q @ et.t() + P["decoder_ob.b"], all_t[:, 2])
        q = conv_tower_train(et[all_t[:, 0]], et[all_t[:, 2]], P, "rdecoder.", stats)
        loss_r = cross_entropy(q @ h0.t() + P["rdecoder.b"], all_t[:, 1])
    ids = torch.unique(all_t[:, [0, 2]].reshape(-1))
    rs = static_radius(P["radius_static"], c, rmin, rmax)[ids]
    loss_rad = radius_lambda * torch.mean((rs - P["radius_target"][ids]) ** 2)
    if static is not None:
        # hyperbolic_model.py:1039-1064: the angle loss against the TANGENT vectors log_0(evolve_emb)
        loss_st = static_angle_loss(s_emb, [log0(e, c) for e in hist], layer_norm, static["angle"], static["discount"],
                                    static["weight"])
        return loss_e, loss_r, loss_rad, loss_st
    return loss_e, loss_r, loss_rad


def hyp_train_steps(sd, graphs, num_rels, triples, c=0.01, layer_norm=False, gamma=0.15, steps=1, task_weight=0.7,
                    grad_norm=1.0, lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, dtype=torch.float32,
                    decoder="hyperbolic_convtranse", encoder="hyperbolic_uvrgcn", num_bases=100, static=None):
    """Like regcn_train_steps for the hyperbolic model.  Returns per-step dicts {losses (e, r, static, radius),
    grad_norm, grads, params}."""
    P = {}
    for k, v in sd.items():
        if not v.is_floating_point() or k == "rgcn.rel_emb":
            continue
        t = v.detach().clone().to(dtype)
        if "running_" not in k and k not in ("c", "log_c", "radius_target"):
            t.requires_grad_(True)
        P[k] = t
    m = {k: torch.zeros_like(v) for k, v in P.items() if v.requires_grad}
    vv = {k: torch.zeros_like(v) for k, v in P.items() if v.requires_grad}
    log = []
    for step in range(1, steps + 1):
        stats = {}
        ls = hyp_train_losses(P, graphs, num_rels, triples, c, layer_norm, gamma, stats, decoder=decoder,
                              encoder=encoder, num_bases=num_bases, static=static)
        le, lrel, lrad = ls[:3]
        lst = ls[3] if static is not None else torch.zeros(())
        loss = task_weight * le + (1 - task_weight) * lrel + lrad + lst
        names = [k for k, v in P.items() if v.requires_grad]
        gs = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
        grads = {k: g for k, g in zip(names, gs) if g is not None}
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
        coef = min(1.0, grad_norm / (float(total) + 1e-6))
        with torch.no_grad():
            for k, g in grads.items():
                p = P[k]
                gg = g * coef + weight_decay * p
                m[k] = betas[0] * m[k] + (1 - betas[0]) * gg
                vv[k] = betas[1] * vv[k] + (1 - betas[1]) * gg * gg
                p -= (lr / (1 - betas[0] ** step)) * m[k] / (vv[k].sqrt() / math.sqrt(1 - betas[1] ** step) + eps)
            for k, s in stats.items():
                P[k] = s
        log.append({"losses": (float(le.detach()), float(lrel.detach()), float(lst.detach()), float(lrad.detach())),
                    "grad_norm": float(total), "grads": {k: g.detach().clone() for k, g in grads.items()},
                    "params": {k: P[k].detach().clone() for k in grads}})
    return log


# --------------------------------------------------------------------------------------------------------------
# TF32 operand split of the tensor-core contractions (not a reference function: the 3xTF32 GEMMs feed every fp32
# operand x as hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi); PTX cvt.rna.tf32.f32 = round to nearest, ties away
# from zero, to 11 significant bits).  Written with frexp / floor in float64 -- independently of the bit trick the
# kernels use -- and valid for normal fp32 values.
# --------------------------------------------------------------------------------------------------------------
def rna_tf32(x):
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    m, e = np.frexp(np.abs(x))                      # |x| = m * 2^e, m in [0.5, 1)
    q = np.floor(m * 2048.0 + 0.5) / 2048.0         # 11 significant bits, ties away from zero
    return (np.sign(x) * np.ldexp(q, e)).astype(np.float32)


def split_tf32(x):
    x = np.asarray(x, dtype=np.float32)
    hi = rna_tf32(x)
    lo = rna_tf32((x.astype(np.float32) - hi).astype(np.float32))
    return hi, lo
