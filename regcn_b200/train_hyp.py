"""Training-mode path of HyperbolicRecurrentRGCN (second half): `get_loss` with gradients for the
hyperbolic_uvrgcn encoder + hyperbolic_convtranse decoder (hyperbolic_model.py:722-890, 941-1088) -- the configuration of
the reference's only published run (hyperbolic_src/train.log).

Same construction as regcn_b200/train.py: torch.autograd is the tape, every node's forward and backward is a kernel.
The Poincare row maps (exp_0 with its projection, log_0, project_to_ball, the tangent normalisation) are radial maps
y = s(|x|) x and share one backward kernel (`regcn_radial_bwd`); the radius-difference edge weights
exp(-gamma |rho_u - rho_v|) (hyperbolic_layers.py:232-234) are differentiated w.r.t. the messages AND the radii."""
import torch

from . import _lib, ops
from ._lib import call, ptr
from . import train as T

F32 = torch.float32
LOG0, EXP0, PROJECT, TNORM = 0, 1, 2, 3
_ROWMAP = {LOG0: ops.ROW_LOG0, EXP0: ops.ROW_EXP0, PROJECT: ops.ROW_PROJECT, TNORM: ops.ROW_TANGENT_NORMALIZE}


class _Radial(torch.autograd.Function):
    """log_0 / exp_0 / project_to_ball / exp_0(normalize(log_0 .))   (hyperbolic_ops.py:38-116)."""

    @staticmethod
    def forward(ctx, x, mode, c):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.mode, ctx.c = mode, float(c)
        return ops.row_map(x, _ROWMAP[mode], c=c)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("regcn_radial_bwd", ptr(x), ptr(dy.contiguous()), ptr(dx), x.shape[0], x.shape[1], ctx.mode, ctx.c)
        return dx, None, None


class _Radius(torch.autograd.Function):
    """get_radius (hyperbolic_ops.py:206)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        rho = torch.empty(x.shape[0], device=x.device, dtype=F32)
        call("regcn_row_radius", ptr(x), x.shape[0], x.shape[1], ptr(rho))
        return rho

    @staticmethod
    def backward(ctx, drho):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("regcn_row_radius_bwd", ptr(x), ptr(drho.contiguous()), x.shape[0], x.shape[1], ptr(dx))
        return dx


class _ApplyRadius(torch.autograd.Function):
    """apply_radius (hyperbolic_ops.py:222-233)."""

    @staticmethod
    def forward(ctx, x, r, c):
        x, r = x.contiguous(), r.contiguous()
        ctx.save_for_backward(x, r)
        ctx.c = float(c)
        y = torch.empty_like(x)
        call("regcn_apply_radius", ptr(x), ptr(r), x.shape[0], x.shape[1], ctx.c, ptr(y))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, r = ctx.saved_tensors
        dx = torch.empty_like(x)
        dr = torch.empty_like(r)
        call("regcn_apply_radius_bwd", ptr(x), ptr(r), ptr(dy.contiguous()), x.shape[0], x.shape[1], ctx.c, ptr(dx), ptr(dr))
        return dx, dr, None


class _Eltwise(torch.autograd.Function):
    """op 0: clamp(x, -lim, lim); op 1: 0.9 tanh(x) + 0.1 x."""

    @staticmethod
    def forward(ctx, x, op, lim):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.op, ctx.lim = int(op), float(lim)
        y = torch.empty_like(x)
        call("regcn_eltwise_fwd", ptr(x), ptr(y), x.numel(), ctx.op, ctx.lim)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("regcn_eltwise_bwd", ptr(x), ptr(dy.contiguous()), ptr(dx), x.numel(), ctx.op, ctx.lim)
        return dx, None, None


class _RadiusCombine(torch.autograd.Function):
    """_static_radius (hyperbolic_model.py:715-720) [+ the scalar part of TemporalRadiusEvolution, ops:408-424]."""

    @staticmethod
    def forward(ctx, raw, dyn, delta, rmin, rmax, c, beta, eps_r):
        raw = raw.contiguous()
        has = dyn is not None
        dyn = dyn.contiguous() if has else None
        delta = delta.contiguous() if has else None
        ctx.save_for_backward(raw, delta if has else raw)
        ctx.cfg = (float(rmin), float(rmax), float(c), float(beta), float(eps_r), has)
        out = torch.empty_like(raw)
        call("regcn_radius_combine", ptr(raw), ptr(dyn), ptr(delta), raw.shape[0], float(rmin), float(rmax), float(c),
             float(beta), float(eps_r), ptr(out))
        return out

    @staticmethod
    def backward(ctx, g):
        raw, delta = ctx.saved_tensors
        rmin, rmax, c, beta, eps_r, has = ctx.cfg
        draw = torch.empty_like(raw)
        ddyn = torch.empty_like(raw) if has else None
        ddelta = torch.empty_like(raw) if has else None
        call("regcn_radius_combine_bwd", ptr(raw), ptr(delta) if has else None, ptr(g.contiguous()), raw.shape[0], rmin, rmax,
             c, beta, eps_r, ptr(draw), ptr(ddyn), ptr(ddelta))
        return draw, ddyn, ddelta, None, None, None, None, None


class _RowDot(torch.autograd.Function):
    """radius_mlp = nn.Linear(d, 1) (hyperbolic_ops.py:390-392,407): delta[n] = <t[n], w> + b."""

    @staticmethod
    def forward(ctx, t, w, b):
        t, w, b = t.contiguous(), w.contiguous(), b.contiguous()
        ctx.save_for_backward(t, w)
        out = torch.empty(t.shape[0], device=t.device, dtype=F32)
        call("regcn_row_dot", ptr(t), ptr(w), ptr(b), t.shape[0], t.shape[1], ptr(out))
        return out

    @staticmethod
    def backward(ctx, dout):
        t, w = ctx.saved_tensors
        dout = dout.contiguous()
        dt = torch.empty_like(t)
        scaled = torch.empty_like(t)
        call("regcn_row_dot_bwd", ptr(t), ptr(w), ptr(dout), t.shape[0], t.shape[1], ptr(dt), ptr(scaled))
        dw = T._col_sum(scaled).view_as(w)
        db = T._col_sum(dout.view(-1, 1)).view(1)
        return dt, dw, db


def _src_index(g):
    """CSR positions grouped by their source entity (the by-source pass of the radius gradient)."""
    idx = getattr(g, "_src_idx", None)
    if idx is None:
        rp, perm, _ = T._group(g.src_sorted[:g.num_edges].contiguous(), g.num_nodes)
        idx = g._src_idx = (rp, perm)
    return idx


class _HypAggregate(torch.autograd.Function):
    """agg[v] = norm[v] sum_{(u,r)->v} exp(-gamma |rho_u - rho_v|) (ht[u] + rel[r])   (hyperbolic_layers.py:222-240)."""

    @staticmethod
    def forward(ctx, ht, rel, rho, g, gamma):
        ht, rel, rho = ht.contiguous(), rel.contiguous(), rho.contiguous()
        ctx.save_for_backward(ht, rel, rho)
        ctx.g, ctx.gamma = g, float(gamma)
        return ops.union_aggregate(ht, rel, g, radius=rho, gamma=gamma)

    @staticmethod
    def backward(ctx, dagg):
        ht, rel, rho = ctx.saved_tensors
        g, gamma = ctx.g, ctx.gamma
        dagg = dagg.contiguous()
        N, d = ht.shape
        type_rowptr, type_src, type_dst = T._block_index(g)
        dht = T._gather_sum(dagg, d, g.norm, g.rowptr, g.src_sorted, N, d, rho=rho, gamma=gamma)
        drel = T._gather_sum(dagg, d, g.norm, type_rowptr, type_dst, 2 * g.num_rels, d, rho=rho, gamma=gamma,
                             partner=type_src)
        s_edge = torch.empty(max(g.num_edges, 1), device=ht.device, dtype=F32)
        drho = torch.empty(N, device=ht.device, dtype=F32)
        call("regcn_edge_radius_grad", ptr(ht), ptr(rel), ptr(dagg), ptr(g.rowptr), ptr(g.src_sorted), ptr(g.etype_sorted),
             ptr(g.norm), ptr(rho), gamma, N, d, ptr(s_edge), ptr(drho))
        rp, perm = _src_index(g)
        call("regcn_edge_scalar_gather", ptr(s_edge), ptr(rp), ptr(perm), N, ptr(drho), 1)
        return dht, drel, drho, None, None


class _SelectAdd(torch.autograd.Function):
    """P + where(indeg > 0, L[:, :d], L[:, d:])   (hyperbolic_layers.py:273-283, 302-308)."""

    @staticmethod
    def forward(ctx, P, L, g):
        ctx.g = g
        out, _, _ = ops.union_combine(P.contiguous(), L.contiguous(), g.indeg, act=0)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        N, d = dout.shape
        dP = torch.empty_like(dout)
        dL = torch.empty((N, 2 * d), device=dout.device, dtype=F32)
        call("regcn_union_combine_bwd", None, ptr(dout), ptr(ctx.g.indeg), N, d, 0.0, ptr(dP), ptr(dL))
        return dP, dL, None


radial = _Radial.apply
radius = _Radius.apply
apply_radius = _ApplyRadius.apply
eltwise = _Eltwise.apply
row_dot = _RowDot.apply


def static_radius(model):
    return _RadiusCombine.apply(model.radius_static, None, None, model.radius_min, model.radius_max, model._c_float, 1.0, 0.0)


def hyp_union_layer(layer, g, h_in, h0, c, training):
    """HyperbolicUnionRGCNLayer.forward (hyperbolic_layers.py:262-323; self_loop, no skip connection)."""
    p = float(layer.dropout.p) if (layer.dropout is not None and training) else 0.0
    ht = radial(h_in, LOG0, c)
    rho = radius(h_in)
    agg = _HypAggregate.apply(ht, h0, rho, g, float(layer.radius_msg_gamma))
    P = eltwise(T.linear(agg, layer.weight_neighbor, None, True), 0, 10.0)
    L = T.linear(ht, T.loop_cat(layer), None, True)
    t = eltwise(_SelectAdd.apply(P, L, g), 0, 10.0)
    t = T._RReluDrop.apply(t, p)
    return radial(t, EXP0, c)


class _LorentzAggregate(torch.autograd.Function):
    """LorentzRGCNLayer message passing (hyperbolic_layers.py:589-625, 665-672): per-edge blockdiag(W[r]) ht[u] + rel[r]
    -> exp_0 -> to_lorentz, per-node Lorentz centroid -> to_poincare -> log_0 -> clamp.  Any block size dividing d (the
    reference clamps num_bases to 2R, :559-561: 2x2 blocks at 100 bases, 10x10 on a 10-relation dataset)."""

    @staticmethod
    def forward(ctx, ht, weight, rel, g, num_bases, c):
        ht, weight, rel = ht.contiguous(), weight.contiguous(), rel.contiguous()
        ctx.save_for_backward(ht, weight, rel)
        ctx.g, ctx.nb, ctx.c = g, int(num_bases), float(c)
        return ops.lorentz_aggregate(ht, weight, rel, g, num_bases, c)

    @staticmethod
    def backward(ctx, gout):
        ht, weight, rel = ctx.saved_tensors
        g = ctx.g
        N, d = ht.shape
        R2 = 2 * g.num_rels
        dev = ht.device
        lib = _lib.load()
        S = lib.regcn_lorentz_bwd_splits()
        type_rowptr, type_src, type_dst = T._block_index(g)
        dht = torch.empty_like(ht)
        part_rel = torch.empty((S, R2 * d), device=dev, dtype=F32)
        part_w = torch.empty((S, R2 * d * (d // ctx.nb)), device=dev, dtype=F32)
        nb = lib.regcn_lorentz_aggregate_bwd_workspace_bytes(N, R2, d)
        ws = T._ws(dev, nb, slot=2)
        call("regcn_lorentz_aggregate_bwd", ptr(ht), ptr(weight), ptr(rel), ptr(gout.contiguous()), ptr(g.rowptr),
             ptr(g.src_sorted), ptr(g.etype_sorted), ptr(g.norm), ptr(type_rowptr), ptr(type_src), ptr(type_dst), N, R2, d,
             ctx.nb, ctx.c, ptr(dht), ptr(part_rel), ptr(part_w), ptr(ws), nb)
        drel = T._col_sum(part_rel).view(R2, d)
        dW = T._col_sum(part_w).view_as(weight)
        return dht, dW, drel, None, None, None


def lorentz_layer(layer, g, h_in, h0, c, training, prev_h=None):
    """LorentzRGCNLayer.forward (hyperbolic_layers.py:627-694; self_loop; skip gate when the cell hands over the
    previous layer's input, :657-662,675-678)."""
    p = float(layer.dropout.p) if (layer.dropout is not None and training) else 0.0
    ht = radial(h_in, LOG0, c)
    agg = _LorentzAggregate.apply(ht, layer.weight, h0, g, layer.num_bases, c)
    L = T.linear(ht, T.loop_cat(layer), None, True)
    t = _SelectAdd.apply(agg, L, g)
    if layer.skip_connect and prev_h is not None:
        pt = radial(prev_h, LOG0, c)
        # sigmoid(pt W_s + b) * t + (1 - sigmoid) * pt: the time-gate node without its normalisation
        t = T.time_gate(T.linear(pt, layer.skip_weight, None, True), layer.skip_bias, t, pt, False)
    t = eltwise(t, 0, 10.0)
    t = T._RReluDrop.apply(t, p)
    return radial(t, EXP0, c)


def hyp_evolve(model, g_list, static_graph=None):
    """HyperbolicRecurrentRGCN.forward with the tape on (hyperbolic_model.py:762-890).  Returns (hist, h_0, static_emb)."""
    # --skip-connect: HyperbolicRGCNCell never passes prev_h (hyperbolic_src/hyperbolic_model.py:152), so the flag is
    # inert for hyperbolic_uvrgcn; LorentzRGCNCell does (hyperbolic_src/hyperbolic_layers.py:737-740)
    if model.encoder_name not in ("hyperbolic_uvrgcn", "lgcn") or any(not l.self_loop for l in model.rgcn.layers):
        raise NotImplementedError("regcn_b200.train_hyp: hyperbolic_uvrgcn / lgcn with self_loop")
    if model.encoder_name == "lgcn" and any(model.h_dim % l.num_bases for l in model.rgcn.layers):
        raise ValueError("regcn_b200.train_hyp: lgcn needs num_bases (clamped to 2R) to divide h_dim")
    layer_fn = hyp_union_layer if model.encoder_name == "hyperbolic_uvrgcn" else lorentz_layer
    c = model._c_float
    cell = model.relation_gru
    static_emb = None
    if getattr(model, "use_static", False) and static_graph is not None:
        static_emb = init = T.static_embedding(model, static_graph)
    else:
        init = T.normalize(model.dynamic_emb) if model.layer_norm else model.dynamic_emb
    h = radial(init, EXP0, c)
    rs = static_radius(model)
    h = apply_radius(h, rs, c)
    tre = model.temporal_radius_evolution
    h0 = None
    hist = []
    for i, g in enumerate(g_list):
        ht = radial(h, LOG0, c)
        x_mean = T.rel_mean_pool(ht, g)
        x_cat = torch.cat((model.emb_rel, x_mean), dim=1)
        hprev = model.emb_rel if i == 0 else h0
        gi = T.linear(x_cat, cell.weight_ih, cell.bias_ih)
        gh = T.linear(hprev, cell.weight_hh, cell.bias_hh)
        h0 = T.gru_gate(gi, gh, hprev, model.layer_norm)
        cur = h
        prev_in = None
        for layer in model.rgcn.layers:
            if model.encoder_name == "lgcn":
                cur, prev_in = lorentz_layer(layer, g, cur, h0, c, model.training, prev_in), cur
            else:
                cur = layer_fn(layer, g, cur, h0, c, model.training)
        cur = radial(cur, PROJECT, c)
        if model.layer_norm:
            cur = radial(cur, TNORM, c)
        ct = eltwise(radial(cur, LOG0, c), 0, 10.0)
        pt = eltwise(radial(h, LOG0, c), 0, 10.0)
        G = T.linear(pt, model.time_gate_weight, None, True)
        nt = T.time_gate(G, model.time_gate_bias, ct, pt, False)
        h = radial(radial(nt, EXP0, c), PROJECT, c)
        if model.use_residual_evolution:
            t = radial(h, LOG0, c)
            delta = row_dot(t, tre.radius_mlp.weight, tre.radius_mlp.bias)
            r_new = _RadiusCombine.apply(model.radius_static, radius(h), delta, model.radius_min, model.radius_max, c,
                                         float(tre.anchor_beta), float(tre.epsilon))
            h = apply_radius(h, r_new, c)
        else:
            h = apply_radius(h, rs, c)
        hist.append(h)
    return hist, h0, static_emb


class _RadiusLoss(torch.autograd.Function):
    """loss_radius = lambda * mse(static_radius[ids], radius_target[ids])   (hyperbolic_model.py:1066-1073)."""

    @staticmethod
    def forward(ctx, raw, target, ids, rmin, rmax, c, lam):
        raw, target, ids = raw.contiguous(), target.contiguous(), ids.contiguous()
        n = int(ids.shape[0])
        term = torch.empty((max(n, 1), 1), device=raw.device, dtype=F32)
        call("regcn_radius_mse", ptr(raw), ptr(target), ptr(ids), n, float(rmin), float(rmax), float(c), float(lam), ptr(term))
        ctx.save_for_backward(raw, target, ids)
        ctx.cfg = (float(rmin), float(rmax), float(c), float(lam))
        return T._col_sum(term) if n else torch.zeros(1, device=raw.device)

    @staticmethod
    def backward(ctx, g):
        raw, target, ids = ctx.saved_tensors
        rmin, rmax, c, lam = ctx.cfg
        draw = torch.zeros_like(raw)
        call("regcn_radius_mse_bwd", ptr(raw), ptr(target), ptr(ids), int(ids.shape[0]), rmin, rmax, c, lam,
             ptr(g.contiguous().view(-1)), ptr(draw))
        return draw, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------ distance decoders
class _GatherRows(torch.autograd.Function):
    """out[b] = table[idx[b]]   (entity / relation lookups of the decoders, hyperbolic_decoder.py:744-746)."""

    @staticmethod
    def forward(ctx, table, idx32):
        table, idx32 = table.contiguous(), idx32.contiguous()
        P, d = int(idx32.shape[0]), table.shape[1]
        out = torch.empty((P, d), device=table.device, dtype=F32)
        call("regcn_gather_rows2", ptr(table), None, ptr(idx32), P, d, ptr(out), None)
        ctx.save_for_backward(idx32)
        ctx.nrows = table.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx32,) = ctx.saved_tensors
        dout = dout.contiguous()
        rp, perm, _ = T._group(idx32, ctx.nrows)
        d = dout.shape[1]
        if d <= 256:
            return T._gather_sum(dout, d, None, rp, perm, ctx.nrows, d), None
        # wider rows (AttH's (2R, 2d) attention table): column blocks of at most 256 through the same kernel
        out = torch.empty((ctx.nrows, d), device=dout.device, dtype=F32)
        nblk = (d + 255) // 256
        step = ((d + nblk - 1) // nblk + 3) // 4 * 4
        for c0 in range(0, d, step):
            w = min(step, d - c0)
            call("regcn_csr_gather_sum", dout.data_ptr() + 4 * c0, d, None, None, ptr(rp), ptr(perm), ctx.nrows, w, 0,
                 out.data_ptr() + 4 * c0, d, 0, None, 0.0, None)
        return out, None


class _Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = x.contiguous(), y.contiguous()
        ctx.save_for_backward(x, y)
        z = torch.empty_like(x)
        call("regcn_eltwise_mul", ptr(x), ptr(y), ptr(z), x.numel())
        return z

    @staticmethod
    def backward(ctx, dz):
        x, y = ctx.saved_tensors
        dz = dz.contiguous()
        dx, dy = torch.empty_like(x), torch.empty_like(x)
        call("regcn_eltwise_mul", ptr(dz), ptr(y), ptr(dx), x.numel())
        call("regcn_eltwise_mul", ptr(dz), ptr(x), ptr(dy), x.numel())
        return dx, dy


class _Mobius(torch.autograd.Function):
    """mobius_add before its projection (hyperbolic_ops.py:135-142); follow with radial(., PROJECT)."""

    @staticmethod
    def forward(ctx, x, y, c):
        x, y = x.contiguous(), y.contiguous()
        ctx.save_for_backward(x, y)
        ctx.c = float(c)
        z = torch.empty_like(x)
        call("regcn_mobius_fwd", ptr(x), ptr(y), x.shape[0], x.shape[1], ctx.c, ptr(z))
        return z

    @staticmethod
    def backward(ctx, dz):
        x, y = ctx.saved_tensors
        dx, dy = torch.empty_like(x), torch.empty_like(x)
        call("regcn_mobius_bwd", ptr(x), ptr(y), ptr(dz.contiguous()), x.shape[0], x.shape[1], ctx.c, ptr(dx), ptr(dy))
        return dx, dy, None


class _Dropout(torch.autograd.Function):
    """nn.Dropout in train mode (mask recovered from the output)."""

    @staticmethod
    def forward(ctx, x, p):
        out = x.contiguous().clone()
        call("regcn_dropout", ptr(out), out.numel(), float(p), T._next_seed())
        ctx.p = float(p)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        M, d = out.shape
        dx = torch.empty_like(out)
        call("regcn_bn_bwd_apply", ptr(dout.contiguous()), ptr(out), None, M, 1, d, 2, 1.0 / (1.0 - ctx.p), None, None, None,
             None, None, None, 0, 1.0, ptr(dx))
        return dx, None


def dropout(x, p, training):
    return _Dropout.apply(x, p) if (training and p > 0) else x


class _HypDistCE(torch.autograd.Function):
    """mean_b CrossEntropy over scale (margin - |(-q_b)(+)_c e_n|^2) + bias_n   (_chunked_hyperbolic_ce_loss,
    hyperbolic_decoder.py:182-307, proxy-distance branch) in the <q,e>, |q|^2, |e|^2 form: one GEMM for the dots, the
    score epilogue in place, and in the backward the three partial derivatives per (query, candidate) feed two GEMMs."""

    @staticmethod
    def forward(ctx, q, cand, bias, scale, margin, triples, target_col, c, curv_raw=None, curv_cfg=None):
        q, cand = q.contiguous(), cand.contiguous()
        B, d = q.shape
        N = cand.shape[0]
        Np = T._pad4(N)
        dev = q.device
        sm = torch.stack((scale.detach().reshape(()), margin.detach().reshape(()))).float().contiguous()
        x2, y2 = ops.row_sumsq(q), ops.row_sumsq(cand)
        S = torch.empty((B, Np), device=dev, dtype=F32)
        T._mm(T._split(q), T._split(cand), B, N, d, out=S, ldc=Np)
        # per-query curvature (--plus-relation-specific-curvature): true-distance branch of the score
        row_c = None
        if curv_raw is not None:
            num_rel, cmax = curv_cfg
            row_c = ops.rel_curvature(curv_raw, triples, num_rel, c, cmax)
        ctx.row_c, ctx.curv_cfg = row_c, curv_cfg
        ctx.curv_raw = curv_raw.detach().contiguous() if curv_raw is not None else None
        call("regcn_hyp_score_epilogue", ptr(S), Np, B, N, ptr(x2), ptr(y2), ptr(bias.contiguous()) if bias is not None else None,
             None, float(c), ptr(sm), ptr(row_c))
        ce = torch.empty(B, device=dev, dtype=F32)
        lse = torch.empty(B, device=dev, dtype=F32)
        loss = torch.empty(1, device=dev, dtype=F32)
        call("regcn_ce_lse_rows", ptr(S), Np, B, N, ptr(triples), target_col, ptr(ce), ptr(lse), ptr(loss))
        ctx.save_for_backward(q, cand, triples, lse, x2, y2, sm)
        ctx.S, ctx.target_col, ctx.c, ctx.has_bias = S, target_col, float(c), bias is not None
        return loss

    @staticmethod
    def backward(ctx, gloss):
        q, cand, triples, lse, x2, y2, sm = ctx.saved_tensors
        S, ctx.S = ctx.S, None
        B, d = q.shape
        N, Np = cand.shape[0], S.shape[1]
        dev = q.device
        call("regcn_softmax_grad_rows", ptr(S), Np, B, N, ptr(triples), ctx.target_col, ptr(lse),
             ptr(gloss.contiguous().view(-1)))                                    # S <- dS
        qs, cs = T._split(q), T._split(cand)
        D = torch.empty((B, Np), device=dev, dtype=F32)
        T._mm(qs, cs, B, N, d, out=D, ldc=Np)                                   # the dots again (cheaper than keeping them)
        H = torch.empty((B, Np), device=dev, dtype=F32)
        gx, gs, gm = (torch.empty(B, device=dev, dtype=F32) for _ in range(3))
        draw = None
        if ctx.row_c is not None:
            gc = torch.empty(B, device=dev, dtype=F32)
            call("regcn_hyp_truedist_grad", ptr(D), ptr(S), ptr(H), Np, B, N, ptr(x2), ptr(y2), ptr(ctx.row_c), ptr(sm), ptr(gx),
                 ptr(gs), ptr(gm), ptr(gc))
            num_rel, cmax = ctx.curv_cfg
            draw_q = torch.empty(B, device=dev, dtype=F32)
            base = torch.empty(B, device=dev, dtype=torch.int32)
            call("regcn_rel_curvature_bwd", ptr(ctx.curv_raw), ptr(triples), B, int(num_rel), ctx.c,
                 float(cmax) if cmax is not None else 0.0, ptr(gc), ptr(draw_q), ptr(base))
            rp, perm, _ = T._group(base, int(ctx.curv_raw.shape[0]))
            draw = torch.empty_like(ctx.curv_raw)
            call("regcn_edge_scalar_gather", ptr(draw_q), ptr(rp), ptr(perm), int(draw.shape[0]), ptr(draw), 0)
        else:
            call("regcn_hyp_dist_grad", ptr(D), ptr(S), ptr(H), Np, B, N, ptr(x2), ptr(y2), ctx.c, ptr(sm), ptr(gx), ptr(gs),
                 ptr(gm))
        Ds = T._split(D)
        dq = T._mm(Ds, cs, B, d, N, b_mn=True).contiguous()
        call("regcn_row_axpy", ptr(q), ptr(gx), 2.0, B, d, ptr(dq))
        dcand = T._mm(Ds, qs, N, d, B, a_mn=True, b_mn=True).contiguous()
        gy = T._col_sum(H[:, :N])
        call("regcn_row_axpy", ptr(cand), ptr(gy), 2.0, N, d, ptr(dcand))
        dbias = T._col_sum(S[:, :N]) if ctx.has_bias else None
        dscale = T._col_sum(gs.view(B, 1)).view(())
        dmargin = T._col_sum(gm.view(B, 1)).view(())
        return dq, dcand, dbias, dscale, dmargin, None, None, None, draw, None


def _curv_args(dec):
    """(rel_curvature_raw, (num_relations, rel_curvature_max)) or (None, None)."""
    raw = getattr(dec, "rel_curvature_raw", None)
    return (raw, (dec.num_relations, dec.rel_curvature_max)) if raw is not None else (None, None)


def murp_ent_loss(dec, pre, r_emb, all_t, c, training):
    """HyperbolicMuRP.loss (hyperbolic_decoder.py:781-817)."""
    s32 = all_t[:, 0].to(torch.int32).contiguous()
    r32 = all_t[:, 1].to(torch.int32).contiguous()
    o32 = all_t[:, 2].to(torch.int32).contiguous()
    s_emb = radial(_GatherRows.apply(pre, s32), PROJECT, c)
    rot = _GatherRows.apply(T.linear(r_emb, dec.rot_proj.weight, dec.rot_proj.bias), r32)
    st = dropout(radial(s_emb, LOG0, c), float(dec.dropout.p), training)
    rs = radial(radial(_Mul.apply(rot, st), EXP0, c), PROJECT, c)
    tr = _GatherRows.apply(T.linear(r_emb, dec.trans_proj.weight, dec.trans_proj.bias), r32)
    tr = radial(radial(tr, EXP0, c), PROJECT, c)
    q = radial(_Mobius.apply(rs, tr, c), PROJECT, c)
    scale = torch.nn.functional.softplus(dec.score_scale_raw) + 1e-6          # two scalars: host-side glue
    return _HypDistCE.apply(q, pre, dec.entity_bias, scale, dec.score_margin, all_t, 2, c, *_curv_args(dec))


def murp_rel_loss(rdec, pre, r_emb, all_t, c, training):
    """HyperbolicMuRPRel.loss (hyperbolic_decoder.py:897-928)."""
    s32 = all_t[:, 0].to(torch.int32).contiguous()
    r32 = all_t[:, 1].to(torch.int32).contiguous()
    o32 = all_t[:, 2].to(torch.int32).contiguous()
    p = float(rdec.dropout.p)
    st = dropout(radial(_GatherRows.apply(pre, s32), LOG0, c), p, training)
    ot = dropout(radial(_GatherRows.apply(pre, o32), LOG0, c), p, training)
    q_tan = T.linear(torch.cat((st, ot), dim=1), torch.cat((rdec.W_s, rdec.W_o), dim=0), None, True)
    q = radial(q_tan, EXP0, c)
    one = torch.ones((), device=pre.device)
    return _HypDistCE.apply(q, radial(r_emb, EXP0, c), rdec.rel_bias, one, one * 0.0, all_t, 1, c)


def murp_losses(model, pre, r_emb, all_t):
    """HyperbolicMuRP.loss / HyperbolicMuRPRel.loss (hyperbolic_decoder.py:781-817, 897-928)."""
    zero = torch.zeros(1, device=pre.device)
    c = model._c_float
    return (murp_ent_loss(model.decoder_ob, pre, r_emb, all_t, c, model.training) if model.entity_prediction else zero,
            murp_rel_loss(model.rdecoder, pre, r_emb, all_t, c, model.training) if model.relation_prediction else zero)


class _Givens(torch.autograd.Function):
    """Givens rotation (mode 0) / reflection (mode 1); ang (B, d/2) or a shared (d/2,) vector."""

    @staticmethod
    def forward(ctx, x, ang, mode):
        x, ang = x.contiguous(), ang.contiguous()
        B, d = x.shape
        ctx.bcast = int(ang.dim() == 1)
        ctx.mode = int(mode)
        ctx.save_for_backward(x, ang)
        y = torch.empty_like(x)
        call("regcn_givens_fwd", ptr(x), ptr(ang), ctx.bcast, B, d, ctx.mode, ptr(y))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, ang = ctx.saved_tensors
        B, d = x.shape
        dx = torch.empty_like(x)
        dang = torch.empty((B, d // 2), device=x.device, dtype=F32)
        call("regcn_givens_bwd", ptr(x), ptr(ang), ptr(dy.contiguous()), ctx.bcast, B, d, ctx.mode, ptr(dx), ptr(dang))
        return dx, (T._col_sum(dang) if ctx.bcast else dang), None


class _Add(torch.autograd.Function):
    """a + b (the residual of RotH's tangent MLP) on the row-axpy kernel."""

    @staticmethod
    def forward(ctx, a, b):
        out = b.contiguous().clone()
        ones = torch.ones(a.shape[0], device=a.device, dtype=F32)
        call("regcn_row_axpy", ptr(a.contiguous()), ptr(ones), 1.0, a.shape[0], a.shape[1], ptr(out))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def _reshape_tangent(dec, x):
    """x + fc2(relu(fc1(x)))   (hyperbolic_decoder.py:1028-1030)."""
    h1 = eltwise(T.linear(x, dec.reshape_fc1.weight, dec.reshape_fc1.bias), 3, 0.0)
    return _Add.apply(x, T.linear(h1, dec.reshape_fc2.weight, dec.reshape_fc2.bias))


def roth_ent_loss(dec, pre, r_emb, all_t, c, training):
    """HyperbolicRotH.loss (hyperbolic_decoder.py:1101-1138)."""
    s32 = all_t[:, 0].to(torch.int32).contiguous()
    r32 = all_t[:, 1].to(torch.int32).contiguous()
    o32 = all_t[:, 2].to(torch.int32).contiguous()
    sp = torch.nn.functional.softplus
    st = radial(radial(_GatherRows.apply(pre, s32), PROJECT, c), LOG0, c)
    st = _reshape_tangent(dec, dropout(st, float(dec.dropout.p), training))
    ang = _GatherRows.apply(T.linear(r_emb, dec.rot_proj.weight, dec.rot_proj.bias), r32)
    rs = radial(radial(_Givens.apply(st, ang, 0), EXP0, c), PROJECT, c)
    tr = _GatherRows.apply(T.linear(r_emb, dec.trans_proj.weight, dec.trans_proj.bias), r32)
    tr = radial(radial(tr, EXP0, c), PROJECT, c)
    q = radial(_Mobius.apply(rs, tr, c), PROJECT, c)
    return _HypDistCE.apply(q, pre, dec.entity_bias, sp(dec.score_scale_raw) + 1e-6, dec.score_margin, all_t, 2, c,
                                *_curv_args(dec))


def roth_rel_loss(rdec, pre, r_emb, all_t, c, training):
    """HyperbolicRotHRel.loss (hyperbolic_decoder.py:1249-1280)."""
    s32 = all_t[:, 0].to(torch.int32).contiguous()
    r32 = all_t[:, 1].to(torch.int32).contiguous()
    o32 = all_t[:, 2].to(torch.int32).contiguous()
    sp = torch.nn.functional.softplus
    st = radial(_GatherRows.apply(pre, s32), LOG0, c)
    st = _reshape_tangent(rdec, dropout(st, float(rdec.dropout.p), training))
    rs = eltwise(radial(_Givens.apply(st, rdec.global_rot, 0), EXP0, c), 2, 0.0)          # -exp_0(rot)
    q = radial(_Mobius.apply(rs, _GatherRows.apply(pre, o32), c), PROJECT, c)
    return _HypDistCE.apply(q, radial(r_emb, EXP0, c), rdec.rel_bias, sp(rdec.score_scale_raw) + 1e-6,
                                rdec.score_margin, all_t, 1, c)


def roth_losses(model, pre, r_emb, all_t):
    """HyperbolicRotH.loss / HyperbolicRotHRel.loss (hyperbolic_decoder.py:1101-1138, 1264-1280)."""
    zero = torch.zeros(1, device=pre.device)
    c = model._c_float
    return (roth_ent_loss(model.decoder_ob, pre, r_emb, all_t, c, model.training) if model.entity_prediction else zero,
            roth_rel_loss(model.rdecoder, pre, r_emb, all_t, c, model.training) if model.relation_prediction else zero)


class _AttnMix(torch.autograd.Function):
    """a = sigmoid(<w, u>), mixed = a rot + (1-a) ref   (hyperbolic_decoder.py:1434-1445, 1617-1625)."""

    @staticmethod
    def forward(ctx, w, u, rot, ref):
        w, u, rot, ref = w.contiguous(), u.contiguous(), rot.contiguous(), ref.contiguous()
        B, d = rot.shape
        ctx.bcast = int(w.dim() == 1)
        a = torch.empty(B, device=rot.device, dtype=F32)
        mixed = torch.empty_like(rot)
        call("regcn_attn_mix_fwd", ptr(w), ctx.bcast, ptr(u), ptr(rot), ptr(ref), B, d, ptr(a), ptr(mixed))
        ctx.save_for_backward(w, u, rot, ref, a)
        return mixed

    @staticmethod
    def backward(ctx, g):
        w, u, rot, ref, a = ctx.saved_tensors
        B, d = rot.shape
        dw = torch.empty((B, 2 * d), device=rot.device, dtype=F32)
        du = torch.empty_like(dw)
        drot, dref = torch.empty_like(rot), torch.empty_like(rot)
        call("regcn_attn_mix_bwd", ptr(w), ctx.bcast, ptr(u), ptr(rot), ptr(ref), ptr(a), ptr(g.contiguous()), B, d, ptr(dw),
             ptr(du), ptr(drot), ptr(dref))
        return (T._col_sum(dw) if ctx.bcast else dw), du, drot, dref


def atth_ent_loss(dec, pre, r_emb, all_t, c, training):
    """HyperbolicAttH.loss (hyperbolic_decoder.py:1464-1512)."""
    s32 = all_t[:, 0].to(torch.int32).contiguous()
    r32 = all_t[:, 1].to(torch.int32).contiguous()
    o32 = all_t[:, 2].to(torch.int32).contiguous()
    sp = torch.nn.functional.softplus

    def table(lin):
        return _GatherRows.apply(T.linear(r_emb, lin.weight, lin.bias), r32)

    st = dropout(radial(radial(_GatherRows.apply(pre, s32), PROJECT, c), LOG0, c), float(dec.dropout.p), training)
    rel_r = _GatherRows.apply(r_emb, r32)
    rot = _Givens.apply(st, table(dec.rot_proj), 0)
    ref = _Givens.apply(st, table(dec.ref_proj), 1)
    mixed = _AttnMix.apply(table(dec.attn_proj), torch.cat((st, rel_r), dim=1), rot, ref)
    mh = radial(radial(mixed, EXP0, c), PROJECT, c)
    tr = radial(radial(table(dec.trans_proj), EXP0, c), PROJECT, c)
    q = radial(_Mobius.apply(mh, tr, c), PROJECT, c)
    return _HypDistCE.apply(q, pre, dec.entity_bias, sp(dec.score_scale_raw) + 1e-6, dec.score_margin, all_t, 2, c,
                                *_curv_args(dec))


def atth_rel_loss(rdec, pre, r_emb, all_t, c, training):
    """HyperbolicAttHRel.loss (hyperbolic_decoder.py:1641-1700)."""
    s32 = all_t[:, 0].to(torch.int32).contiguous()
    r32 = all_t[:, 1].to(torch.int32).contiguous()
    o32 = all_t[:, 2].to(torch.int32).contiguous()
    sp = torch.nn.functional.softplus

    def table(lin):
        return _GatherRows.apply(T.linear(r_emb, lin.weight, lin.bias), r32)

    o_emb = _GatherRows.apply(pre, o32)
    st = dropout(radial(_GatherRows.apply(pre, s32), LOG0, c), float(rdec.dropout.p), training)
    ot = radial(o_emb, LOG0, c)
    rot = _Givens.apply(st, rdec.global_rot, 0)
    ref = _Givens.apply(st, rdec.global_ref, 1)
    mixed = _AttnMix.apply(rdec.attn_weight, torch.cat((st, ot), dim=1), rot, ref)
    mh = eltwise(radial(mixed, EXP0, c), 2, 0.0)
    q = radial(_Mobius.apply(mh, o_emb, c), PROJECT, c)
    return _HypDistCE.apply(q, radial(r_emb, EXP0, c), rdec.rel_bias, sp(rdec.score_scale_raw) + 1e-6,
                                rdec.score_margin, all_t, 1, c)


def atth_losses(model, pre, r_emb, all_t):
    """HyperbolicAttH.loss / HyperbolicAttHRel.loss (hyperbolic_decoder.py:1482-1512, 1642-1700)."""
    zero = torch.zeros(1, device=pre.device)
    c = model._c_float
    return (atth_ent_loss(model.decoder_ob, pre, r_emb, all_t, c, model.training) if model.entity_prediction else zero,
            atth_rel_loss(model.rdecoder, pre, r_emb, all_t, c, model.training) if model.relation_prediction else zero)


def hyp_get_loss(model, glist, triples, static_graph=None):
    """hyperbolic_model.py:941-1088 with gradients: (loss_ent, loss_rel, loss_static, loss_radius), each (1,)."""
    _lib.require_device()
    T.begin_step()
    if ops.gemm_impl() != "tc":
        raise RuntimeError("regcn_b200.train_hyp needs the tensor-core GEMM (REGCN_GEMM=tc)")
    if model.decoder_name not in ("hyperbolic_convtranse", "murp", "roth", "atth"):
        raise NotImplementedError(f"regcn_b200.train_hyp: unknown decoder {model.decoder_name!r}")
    dev = model.dynamic_emb.device
    c = model._c_float
    triples = torch.as_tensor(triples).to(dev)
    inverse = triples.flip(1)
    inverse[:, 1] = inverse[:, 1] + model.num_rels
    all_triples = torch.cat([triples, inverse]).contiguous()
    hist, r_emb, static_emb = hyp_evolve(model, glist, static_graph)
    pre = radial(hist[-1], TNORM, c) if model.layer_norm else hist[-1]
    loss_ent = torch.zeros(1, device=dev)
    loss_rel = torch.zeros(1, device=dev)
    loss_static = torch.zeros(1, device=dev)
    if static_emb is not None and model.discount in (0, 1):
        # hyperbolic_model.py:1039-1064: angle loss between the static embedding and the TANGENT vectors log_0(evolve_emb)
        loss_static = T._StaticAngle.apply(static_emb, model.layer_norm, model.angle, model.discount, model.weight,
                                           *[radial(e, LOG0, c) for e in hist])
    ids = torch.unique(all_triples[:, [0, 2]].reshape(-1))
    loss_radius = _RadiusLoss.apply(model.radius_static, model.radius_target, ids, model.radius_min, model.radius_max, c,
                                    float(model.radius_lambda))
    if model.decoder_name in ("murp", "roth", "atth"):
        fn = {"murp": murp_losses, "roth": roth_losses, "atth": atth_losses}[model.decoder_name]
        loss_ent, loss_rel = fn(model, pre, r_emb, all_triples)
        return loss_ent, loss_rel, loss_static, loss_radius
    et = eltwise(radial(pre, LOG0, c), 1, 0.0)                      # 0.9 tanh(log_0 E) + 0.1 log_0 E  (:377-379)
    if model.entity_prediction:
        q = T.conv_tower(model.decoder_ob, et, r_emb, all_triples, 0, 1)
        loss_ent = T.score_ce(q, et, all_triples, 2, model.decoder_ob.b)
    if model.relation_prediction:
        q = T.conv_tower(model.rdecoder, et, et, all_triples, 0, 2)
        loss_rel = T.score_ce(q, r_emb, all_triples, 1, model.rdecoder.b)
    return loss_ent, loss_rel, loss_static, loss_radius
