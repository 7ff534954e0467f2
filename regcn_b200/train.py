"""Training-mode path of RecurrentRGCN: `get_loss` with gradients, src/rrgcn.py:197-223,
for the optimisation step of src/main.py:235-246.

torch.autograd is used as the tape only: every node below is a `torch.autograd.Function` whose forward AND backward
are C-ABI kernel calls (csrc/backward.cu + the forward kernels); no torch arithmetic runs on the path apart from
autograd's own gradient accumulation for tensors with several consumers.  Dense contractions (forward, dX and dW)
go through the tcgen05 3xTF32 GEMM; dW = x^T dy, y = x W and dX = dY W hand their operands to the tensor core in the
natural row-major layout (MN-major tiles, `regcn_gemm_tf32_mn`), so nothing is transposed.  Dropout masks come from a counter-based generator (`manual_seed`), so they cannot
match torch's generator element for element -- parity tests run with dropout 0, statistics tests with dropout on.
"""
import torch

from . import _lib, ops
from ._lib import call, ptr

F32 = torch.float32
I32 = torch.int32

_rng = {"seed": 0x5EED1234, "ctr": 0}


def manual_seed(seed):
    """Seed of the dropout masks (the kernels hash (seed, call counter, element index))."""
    _rng["seed"] = int(seed) & 0xFFFFFFFF
    _rng["ctr"] = 0


def _next_seed():
    _rng["ctr"] += 1
    return (_rng["seed"] * 0x9E3779B1 + _rng["ctr"] * 0x85EBCA6B) & 0xFFFFFFFF


def _pad4(n):
    return (int(n) + 3) // 4 * 4


_ws_cache = {}


def _ws(dev, nbytes, slot=0):
    """Grow-only scratch buffer per (device, slot); kernels on one stream use it one after the other."""
    key = (dev, slot)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), device=dev, dtype=torch.uint8)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------ dense helpers
def _split(x):
    """(hi, lo) TF32 split of a contiguous (M, K) matrix, K % 4 == 0."""
    return ops.split_tf32(x.contiguous())


_wsplit_cache = {}


def _split_w(W):
    """TF32 split of a WEIGHT operand, shared by every use inside one optimisation step: the same matrix is the B
    operand of L snapshots' forward GEMMs and of their dX GEMMs.  Keyed on (storage pointer, version, shape); the entry
    keeps the tensor alive so the pointer cannot be recycled; `begin_step()` drops the table."""
    key = (W.data_ptr(), W._version, tuple(W.shape), W.stride(0))
    hit = _wsplit_cache.get(key)
    if hit is None:
        if len(_wsplit_cache) > 256:          # a caller that never reaches begin_step() (decoder.loss() on its own)
            _wsplit_cache.clear()
        hit = _wsplit_cache[key] = (W, _split(W))
    return hit[1]


def begin_step():
    """Called by the get_loss entry points: weight splits of the previous step are stale after the optimiser update."""
    _wsplit_cache.clear()


def loop_cat(layer):
    """[W_loop | W_evolve] (d, 2d) once per layer and step: one autograd node (and one TF32 split) shared by all
    snapshots.  Only for callers that ran begin_step() (the get_loss entry points)."""
    key = ("loop_cat", id(layer), layer.loop_weight._version, layer.evolve_loop_weight._version)
    hit = _wsplit_cache.get(key)
    if hit is None or hit[0] is not layer.loop_weight:          # the entry pins the parameter: ids cannot be recycled
        hit = _wsplit_cache[key] = (layer.loop_weight, torch.cat((layer.loop_weight, layer.evolve_loop_weight), dim=1))
    return hit[1]


def _auto_split_k(M, N, K):
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if K < 1024 or tiles >= 120:
        return 1
    return max(1, min(32, 148 // tiles, K // 256))


def _mm(a, b, M, N, K, a_mn=False, b_mn=False, bias=None, out=None, ldc=None, lda=None, ldb=None):
    """C (M, N) = op(A) op(B) on the 3xTF32 tcgen05 GEMM; a, b = (hi, lo) pairs in their natural row-major layout:
       a_mn=False: A is (M, K);  a_mn=True: A is given as X (K, M) and the product uses X^T   (dW = x^T dy)
       b_mn=False: B is (N, K) and the product uses B^T;  b_mn=True: B is given as Y (K, N)    (y = x W, dX = dY W)
    MN-major operands go to the tensor core as they lie in memory (csrc/gemm_tc.cu): no transposes anywhere."""
    dev = a[0].device
    if out is None:
        ldc = _pad4(N)
        out = torch.empty((M, ldc), device=dev, dtype=F32)
        out = out[:, :N] if ldc != N else out
    elif ldc is None:
        ldc = out.stride(0)
    lda = a[0].stride(0) if lda is None else lda
    ldb = b[0].stride(0) if ldb is None else ldb
    split_k = _auto_split_k(M, N, K)
    ws, ws_bytes = None, 0
    if split_k > 1:
        ws_bytes = _lib.load().regcn_gemm_tf32_workspace_bytes(M, N, split_k)
        ws = _ws(dev, ws_bytes, slot=1)
    call("regcn_gemm_tf32_mn", a[0].data_ptr(), a[1].data_ptr(), lda, b[0].data_ptr(), b[1].data_ptr(), ldb,
         out.data_ptr(), ldc, M, N, K, int(a_mn), int(b_mn), ptr(bias), 0, 3, split_k, ptr(ws), ws_bytes)
    return out


def _col_sum(x):
    rows, cols = x.shape
    if x.stride(1) != 1:
        x = x.contiguous()
    out = torch.empty(cols, device=x.device, dtype=F32)
    nb = _lib.load().regcn_col_reduce_workspace_bytes(rows, cols)
    ws = _ws(x.device, nb)
    call("regcn_col_sum", x.data_ptr(), x.stride(0), rows, cols, ptr(out), 0, ptr(ws), nb)
    return out


class _Linear(torch.autograd.Function):
    """y = x W (+ b) with W (K, N)  [w_kn=True: torch.mm(x, W)]  or  y = x W^T (+ b) with W (N, K) [F.linear]."""

    @staticmethod
    def forward(ctx, x, W, bias, w_kn):
        x = x.contiguous()
        W = W.contiguous()
        M, K = x.shape
        N = W.shape[1] if w_kn else W.shape[0]
        if K % 4 or N % 4:
            raise ValueError("regcn_b200.train: matrix dimensions must be multiples of 4")
        xs = _split(x)
        y = _mm(xs, _split_w(W), M, N, K, b_mn=w_kn, bias=None if bias is None else bias.contiguous())
        ctx.save_for_backward(x, W)
        ctx.xs = xs                                   # the TF32 split of x is the A operand of dW again
        ctx.w_kn = w_kn
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        xs, ctx.xs = ctx.xs, None
        dy = dy.contiguous()
        M, K = x.shape
        N = dy.shape[1]
        dx = dW = db = None
        dys = _split(dy)
        if ctx.needs_input_grad[0]:
            # w_kn: dx = dy W^T, W (K, N) is already the K-major B operand;  F.linear: dx = dy W with W (N, K) as Y
            dx = _mm(dys, _split_w(W), M, K, N, b_mn=not ctx.w_kn)
        if ctx.needs_input_grad[1]:
            # dW = x^T dy (K, N)  /  dy^T x (N, K): both operands MN-major, reduction over the M rows
            dW = _mm(xs, dys, K, N, M, a_mn=True, b_mn=True) if ctx.w_kn else _mm(dys, xs, N, K, M, a_mn=True, b_mn=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _col_sum(dy)
        return dx, dW, db, None


def linear(x, W, bias=None, w_kn=False):
    return _Linear.apply(x, W, bias, w_kn)


# ------------------------------------------------------------------------------------------------ graph transposes
def _group(keys, nkeys, vals=None):
    n = int(keys.shape[0])
    dev = keys.device
    rowptr = torch.empty(nkeys + 1, device=dev, dtype=I32)
    perm = torch.empty(max(n, 1), device=dev, dtype=I32)
    vout = torch.empty(max(n, 1), device=dev, dtype=I32) if vals is not None else None
    nb = _lib.load().regcn_group_by_key_workspace_bytes(n)
    ws = _ws(dev, nb)
    call("regcn_group_by_key", ptr(keys), n, nkeys, ptr(vals), ptr(rowptr), ptr(perm), ptr(vout), ptr(ws), nb)
    return rowptr, perm, vout


def _train_index(g):
    """Transposed indices of one snapshot for the backward gathers (cached on the graph object):
    edges grouped by relation type (-> destination ids) and relation memberships grouped by entity."""
    idx = getattr(g, "_train_idx", None)
    if idx is None:
        E, R, N = g.num_edges, g.num_rels, g.num_nodes
        type_rowptr, _, type_dst = _group(g.etype[:E].contiguous(), 2 * R, g.dst[:E].contiguous())
        nnz = int(g.n_rel_ents)
        rowid = torch.empty(max(nnz, 1), device=g.device, dtype=I32)
        inv_len = torch.empty(R, device=g.device, dtype=F32)
        call("regcn_expand_rowptr", ptr(g.rel_rowptr), R, nnz, ptr(rowid), ptr(inv_len))
        ent_rowptr, _, ent_rel = _group(g.rel_ents[:nnz].contiguous(), N, rowid[:nnz].contiguous())
        idx = g._train_idx = dict(type_rowptr=type_rowptr, type_dst=type_dst, ent_rowptr=ent_rowptr, ent_rel=ent_rel,
                                  inv_len=inv_len)
    return idx


def _gather_sum(X, ldx, col_w, rowptr, col, nrows, d, col2_off=0, x_ptr=None, rho=None, gamma=0.0, partner=None):
    out = torch.empty((nrows, d), device=X.device, dtype=F32)
    call("regcn_csr_gather_sum", X.data_ptr() if x_ptr is None else x_ptr, ldx, ptr(col_w), None, ptr(rowptr), ptr(col),
         nrows, d, col2_off, ptr(out), d, 0, ptr(rho), float(gamma), ptr(partner))
    return out


class _RelMeanPool(torch.autograd.Function):
    """src/rrgcn.py:161-166."""

    @staticmethod
    def forward(ctx, h, g):
        ctx.g = g
        return ops.rel_mean_pool(h, g)

    @staticmethod
    def backward(ctx, gx):
        g = ctx.g
        ti = _train_index(g)
        gx = gx.contiguous()
        d = gx.shape[1]
        dh = _gather_sum(gx, d, ti["inv_len"], ti["ent_rowptr"], ti["ent_rel"], g.num_nodes, d, col2_off=g.num_rels)
        return dh, None


class _UnionAggregate(torch.autograd.Function):
    """agg[v] = norm[v] sum_{(u,r)->v} (h[u] + rel[r])   (rgcn/layers.py:257-279 before the W_n product)."""

    @staticmethod
    def forward(ctx, h, rel, g):
        ctx.g = g
        return ops.union_aggregate(h.contiguous(), rel.contiguous(), g)

    @staticmethod
    def backward(ctx, dagg):
        g = ctx.g
        ti = _train_index(g)
        dagg = dagg.contiguous()
        d = dagg.shape[1]
        dh = _gather_sum(dagg, d, g.norm, g.rowptr, g.src_sorted, g.num_nodes, d)
        drel = _gather_sum(dagg, d, g.norm, ti["type_rowptr"], ti["type_dst"], 2 * g.num_rels, d)
        return dh, drel, None


class _UnionCombine(torch.autograd.Function):
    """out = dropout_p(rrelu(P + where(indeg>0, L[:, :d], L[:, d:])))   (rgcn/layers.py:241-253)."""

    @staticmethod
    def forward(ctx, P, L, g, p):
        out, _, _ = ops.union_combine(P.contiguous(), L.contiguous(), g.indeg, act=1)
        if p > 0:
            call("regcn_dropout", ptr(out), out.numel(), float(p), _next_seed())
        ctx.g, ctx.p = g, float(p)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        N, d = out.shape
        dP = torch.empty((N, d), device=out.device, dtype=F32)
        dL = torch.empty((N, 2 * d), device=out.device, dtype=F32)
        call("regcn_union_combine_bwd", ptr(out), ptr(dout.contiguous()), ptr(ctx.g.indeg), N, d, ctx.p, ptr(dP), ptr(dL))
        return dP, dL, None, None


class _UnionSum(torch.autograd.Function):
    """P + where(indeg>0, L[:, :d], L[:, d:]) without activation: the node representation in front of a LIVE skip gate
    (rgcn/layers.py:236-245)."""

    @staticmethod
    def forward(ctx, P, L, g):
        out, _, _ = ops.union_combine(P.contiguous(), L.contiguous(), g.indeg, act=0)
        ctx.g = g
        return out

    @staticmethod
    def backward(ctx, dout):
        N, d = dout.shape
        dP = torch.empty((N, d), device=dout.device, dtype=F32)
        dL = torch.empty((N, 2 * d), device=dout.device, dtype=F32)
        call("regcn_union_combine_bwd", None, ptr(dout.contiguous()), ptr(ctx.g.indeg), N, d, 0.0, ptr(dP), ptr(dL))
        return dP, dL, None


class _TimeGate(torch.autograd.Function):
    """h' = s(G+b) [normalize](cur) + (1 - s(G+b)) h   (src/rrgcn.py:176-178)."""

    @staticmethod
    def forward(ctx, G, bias, cur, h, normalize):
        G, bias, cur, h = G.contiguous(), bias.contiguous(), cur.contiguous(), h.contiguous()
        ctx.save_for_backward(G, bias, cur, h)
        ctx.normalize = bool(normalize)
        return ops.time_gate(G, bias, cur, h, normalize)

    @staticmethod
    def backward(ctx, dout):
        G, bias, cur, h = ctx.saved_tensors
        N, d = h.shape
        dG, dcur, dh = (torch.empty((N, d), device=h.device, dtype=F32) for _ in range(3))
        call("regcn_time_gate_bwd", ptr(G), ptr(bias), ptr(cur), ptr(h), ptr(dout.contiguous()), N, d,
             int(ctx.normalize), ptr(dG), ptr(dcur), ptr(dh))
        return dG, _col_sum(dG), dcur, dh, None


class _GRUGate(torch.autograd.Function):
    """nn.GRUCell gates (+ F.normalize) from gi, gh   (src/rrgcn.py:168-174)."""

    @staticmethod
    def forward(ctx, gi, gh, hprev, normalize):
        gi, gh, hprev = gi.contiguous(), gh.contiguous(), hprev.contiguous()
        ctx.save_for_backward(gi, gh, hprev)
        ctx.normalize = bool(normalize)
        return ops.gru_gate(gi, gh, hprev, normalize)

    @staticmethod
    def backward(ctx, dout):
        gi, gh, hprev = ctx.saved_tensors
        M, d = hprev.shape
        dgi = torch.empty((M, 3 * d), device=gi.device, dtype=F32)
        dgh = torch.empty((M, 3 * d), device=gi.device, dtype=F32)
        dh = torch.empty((M, d), device=gi.device, dtype=F32)
        call("regcn_gru_gate_bwd", ptr(gi), ptr(gh), ptr(hprev), ptr(dout.contiguous()), M, d, int(ctx.normalize),
             ptr(dgi), ptr(dgh), ptr(dh))
        return dgi, dgh, dh, None


class _Normalize(torch.autograd.Function):
    """F.normalize rows   (src/rrgcn.py:154,206)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.row_map(x, ops.ROW_NORMALIZE)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("regcn_normalize_bwd", ptr(x), ptr(dy.contiguous()), ptr(dx), x.shape[0], x.shape[1])
        return dx


class _Tanh(torch.autograd.Function):
    """tanh of the entity table   (src/decoder.py:30,79)."""

    @staticmethod
    def forward(ctx, x):
        y = ops.row_map(x.contiguous(), ops.ROW_TANH)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dx = torch.empty_like(y)
        call("regcn_tanh_bwd", ptr(y), ptr(dy.contiguous()), ptr(dx), y.numel())
        return dx


# ------------------------------------------------------------------------------------------------ decoder tower
def _bn_stats(x, B, C, L, bn):
    dev = x.device
    mean = torch.empty(C, device=dev, dtype=F32)
    invstd = torch.empty(C, device=dev, dtype=F32)
    nb = _lib.load().regcn_col_reduce_workspace_bytes(B, C * L)
    ws = _ws(dev, nb)
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    track = bn.track_running_stats and bn.running_mean is not None
    call("regcn_bn_stats", ptr(x), B, C, L, float(bn.eps), mom, ptr(mean), ptr(invstd),
         ptr(bn.running_mean) if track else None, ptr(bn.running_var) if track else None, ptr(ws), nb)
    if track:
        # the kernel wrote the running statistics through raw pointers: bump their version counters, the folded-BatchNorm
        # caches of the inference path (decoder._fold_bn, evaluate._tower_table) key on them
        torch.autograd.graph.increment_version(bn.running_mean)
        torch.autograd.graph.increment_version(bn.running_var)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return mean, invstd


def _bn_bwd(dZ, Z, Y, B, C, L, mask_mode, mask_scale, mean, invstd, gamma, out_src=None, out_mode=0, out_scale=1.0):
    dev = dZ.device
    sdy = torch.empty(C, device=dev, dtype=F32)
    sdyx = torch.empty(C, device=dev, dtype=F32)
    nb = _lib.load().regcn_col_reduce_workspace_bytes(B, C * L)
    ws = _ws(dev, nb)
    call("regcn_bn_bwd_stats", ptr(dZ), ptr(Z), ptr(Y), B, C, L, mask_mode, float(mask_scale), ptr(mean), ptr(invstd),
         ptr(sdy), ptr(sdyx), ptr(ws), nb)
    dX = torch.empty_like(Y)
    call("regcn_bn_bwd_apply", ptr(dZ), ptr(Z), ptr(Y), B, C, L, mask_mode, float(mask_scale), ptr(mean), ptr(invstd),
         ptr(gamma), ptr(sdy), ptr(sdyx), ptr(out_src), out_mode, float(out_scale), ptr(dX))
    return dX, sdyx, sdy          # dX, dgamma, dbeta


class _ConvTower(torch.autograd.Function):
    """The ConvTransE / ConvTransR query tower in train mode (src/decoder.py:35-50, 83-95): stacked gathers -> bn0 ->
    input dropout -> Conv1d(2,C,3) -> bn1 -> relu -> feature dropout -> fc -> hidden dropout -> bn2 -> relu.
    BatchNorm uses batch statistics and updates the running ones (momentum 0.1, unbiased variance)."""

    @staticmethod
    def forward(ctx, first, second, g0, b0, wc, bc, g1, b1, wf, bf, g2, b2, triples, col0, col1, mod):
        dev = first.device
        first, second = first.contiguous(), second.contiguous()
        B = int(triples.shape[0])
        d = first.shape[1]
        C, _, ksz = wc.shape
        if B < 2:
            raise ValueError("regcn_b200.train: BatchNorm in train mode needs more than one query")
        p_in, p_feat, p_hid = float(mod.inp_drop.p), float(mod.feature_map_drop.p), float(mod.hidden_drop.p)
        X0 = torch.empty((B, 2, d), device=dev, dtype=F32)
        call("regcn_dec_gather_stack", ptr(first), ptr(second), ptr(triples), col0, col1, B, d, ptr(X0))
        m0, is0 = _bn_stats(X0, B, 2, d, mod.bn0)
        X1 = torch.empty((B, 2, d), device=dev, dtype=F32)
        Y = torch.empty((B, C, d), device=dev, dtype=F32)
        wc_c, bc_c = wc.contiguous(), bc.contiguous()
        call("regcn_dec_conv_fwd", ptr(X0), B, d, C, ksz, ptr(m0), ptr(is0), ptr(g0.contiguous()), ptr(b0.contiguous()),
             p_in, _next_seed(), ptr(wc_c), ptr(bc_c), ptr(X1), ptr(Y))
        m1, is1 = _bn_stats(Y, B, C, d, mod.bn1)
        Z = torch.empty((B, C * d), device=dev, dtype=F32)
        call("regcn_bn_act_drop", ptr(Y), B, C, d, ptr(m1), ptr(is1), ptr(g1.contiguous()), ptr(b1.contiguous()), 1,
             p_feat, _next_seed(), ptr(Z))
        wf_c = wf.contiguous()
        Zs = _split(Z)
        Fq = _mm(Zs, _split(wf_c), B, d, C * d, bias=bf.contiguous())
        Fq = Fq.contiguous()
        if p_hid > 0:
            call("regcn_dropout", ptr(Fq), Fq.numel(), p_hid, _next_seed())
        m2, is2 = _bn_stats(Fq, B, d, 1, mod.bn2)
        Q = torch.empty((B, d), device=dev, dtype=F32)
        call("regcn_bn_act_drop", ptr(Fq), B, d, 1, ptr(m2), ptr(is2), ptr(g2.contiguous()), ptr(b2.contiguous()), 1,
             0.0, 0, ptr(Q))
        ctx.save_for_backward(X0, X1, Y, Z, Fq, Q, m0, is0, m1, is1, m2, is2, g0, wc_c, g1, wf_c, g2, triples)
        ctx.dims = (B, d, C, ksz, col0, col1, first.shape[0], second.shape[0], p_in, p_feat, p_hid)
        ctx.Zs = Zs
        return Q

    @staticmethod
    def backward(ctx, dQ):
        X0, X1, Y, Z, Fq, Q, m0, is0, m1, is1, m2, is2, g0, wc, g1, wf, g2, triples = ctx.saved_tensors
        B, d, C, ksz, col0, col1, n_first, n_second, p_in, p_feat, p_hid = ctx.dims
        dev = dQ.device
        dQ = dQ.contiguous()
        # bn2 + relu (mask Q > 0), then the hidden dropout in front of it (mask Fq != 0)
        dF, dg2, db2 = _bn_bwd(dQ, Q, Fq, B, d, 1, 1, 1.0, m2, is2, g2.contiguous(), out_src=Fq,
                               out_mode=2 if p_hid > 0 else 0, out_scale=1.0 / (1.0 - p_hid) if p_hid > 0 else 1.0)
        dbf = _col_sum(dF)
        Zs, ctx.Zs = ctx.Zs, None
        dFs = _split(dF)
        dwf = _mm(dFs, Zs, d, C * d, B, a_mn=True, b_mn=True)                  # (d, C d) = dF^T Z
        dZ = _mm(dFs, _split(wf), B, C * d, d, b_mn=True)                      # (B, C d) = dF W_fc
        del Zs
        dZ = dZ.contiguous()
        # bn1 + relu + feature dropout (mask Z > 0, scale 1/(1-p))
        dY, dg1, db1 = _bn_bwd(dZ, Z, Y, B, C, d, 1, 1.0 / (1.0 - p_feat) if p_feat > 0 else 1.0, m1, is1,
                               g1.contiguous())
        nb = _lib.load().regcn_dec_conv_bwd_weight_workspace_bytes(B, C)
        ws = _ws(dev, nb)
        dwb = torch.empty(C * 2 * ksz + C, device=dev, dtype=F32)
        call("regcn_dec_conv_bwd_weight", ptr(dY), ptr(X1), B, d, C, ksz, ptr(dwb), ptr(ws), nb)
        dwc = dwb[: C * 2 * ksz].view(C, 2, ksz)
        dbc = dwb[C * 2 * ksz:]
        dX1 = torch.empty((B, 2, d), device=dev, dtype=F32)
        call("regcn_dec_conv_bwd_input", ptr(dY), ptr(X1), B, d, C, ksz, ptr(wc), p_in, ptr(dX1))
        dX0, dg0, db0 = _bn_bwd(dX1, None, X0, B, 2, d, 0, 1.0, m0, is0, g0.contiguous())
        # scatter back to the table rows: queries grouped by the gathered id
        t32 = triples.to(I32)
        dfirst = dsecond = None
        if ctx.needs_input_grad[0]:
            rp, perm, _ = _group(t32[:, col0].contiguous(), n_first)
            dfirst = _gather_sum(dX0, 2 * d, None, rp, perm, n_first, d)
        if ctx.needs_input_grad[1]:
            rp, perm, _ = _group(t32[:, col1].contiguous(), n_second)
            dsecond = _gather_sum(dX0, 2 * d, None, rp, perm, n_second, d, x_ptr=dX0.data_ptr() + 4 * d)
        return (dfirst, dsecond, dg0, db0, dwc, dbc, dg1, db1, dwf, dbf, dg2, db2, None, None, None, None)


def conv_tower(mod, first, second, triples, col0, col1):
    return _ConvTower.apply(first, second, mod.bn0.weight, mod.bn0.bias, mod.conv1.weight, mod.conv1.bias,
                            mod.bn1.weight, mod.bn1.bias, mod.fc.weight, mod.fc.bias, mod.bn2.weight, mod.bn2.bias,
                            triples, col0, col1, mod)


class _ScoreCE(torch.autograd.Function):
    """loss = mean_b CrossEntropy(q_b . cand^T, target_b)   (src/rrgcn.py:218-223; decoder :96-99 / :51).
    The logits are materialised once (pitch padded to 4), turned into their own gradient in place in the backward,
    and contracted twice: dq = dS cand, dcand = dS^T q."""

    @staticmethod
    def forward(ctx, q, cand, triples, target_col, col_bias=None):
        q, cand = q.contiguous(), cand.contiguous()
        B, d = q.shape
        N = cand.shape[0]
        Np = _pad4(N)
        dev = q.device
        S = torch.empty((B, Np), device=dev, dtype=F32)
        _mm(_split(q), _split(cand), B, N, d, out=S, ldc=Np, bias=None if col_bias is None else col_bias.contiguous())
        ctx.has_bias = col_bias is not None
        ce = torch.empty(B, device=dev, dtype=F32)
        lse = torch.empty(B, device=dev, dtype=F32)
        loss = torch.empty(1, device=dev, dtype=F32)
        call("regcn_ce_lse_rows", ptr(S), Np, B, N, ptr(triples), target_col, ptr(ce), ptr(lse), ptr(loss))
        ctx.save_for_backward(q, cand, triples, lse)
        ctx.S = S
        ctx.target_col = target_col
        return loss

    @staticmethod
    def backward(ctx, gloss):
        q, cand, triples, lse = ctx.saved_tensors
        S = ctx.S
        ctx.S = None
        if S is None:
            raise RuntimeError("regcn_b200.train: the loss graph can be back-propagated once")
        B, d = q.shape
        N = cand.shape[0]
        Np = S.shape[1]
        call("regcn_softmax_grad_rows", ptr(S), Np, B, N, ptr(triples), ctx.target_col, ptr(lse),
             ptr(gloss.contiguous().view(-1)))
        dq = dcand = None
        Ss = _split(S)                                                            # (B, Np) pitch, padding columns zero
        if ctx.needs_input_grad[0]:
            dq = _mm(Ss, _split(cand), B, d, N, b_mn=True)                        # dS cand: cand (N, d) as Y
        if ctx.needs_input_grad[1]:
            dcand = _mm(Ss, _split(q), N, d, B, a_mn=True, b_mn=True)             # dS^T q
        db = _col_sum(S[:, :N]) if ctx.has_bias else None                         # scores + b  (hyperbolic_decoder.py:411)
        return dq, dcand, None, None, db


def score_ce(q, cand, triples, target_col, col_bias=None):
    return _ScoreCE.apply(q, cand, triples, target_col, col_bias)


# ------------------------------------------------------------------------------------------------ static-graph constraint
def _block_index(g):
    """Edges of a graph grouped by relation type with their endpoints (the dW gather of the block layer)."""
    idx = getattr(g, "_block_idx", None)
    if idx is None:
        E, R2 = g.num_edges, 2 * g.num_rels
        et = g.etype[:E].contiguous()
        type_rowptr, _, type_src = _group(et, R2, g.src[:E].contiguous())
        _, _, type_dst = _group(et, R2, g.dst[:E].contiguous())
        idx = g._block_idx = (type_rowptr, type_src, type_dst)
    return idx


class _BlockAggregate(torch.autograd.Function):
    """agg[v] = norm[v] sum_{(u,r)->v} blockdiag(W[r]) h[u]   (rgcn/layers.py:167-179)."""

    @staticmethod
    def forward(ctx, h, weight, g, num_bases, d_out):
        h, weight = h.contiguous(), weight.contiguous()
        ctx.save_for_backward(h, weight)
        ctx.g, ctx.nb, ctx.d_out = g, int(num_bases), int(d_out)
        return ops.block_aggregate(h, weight, g, num_bases, d_out)

    @staticmethod
    def backward(ctx, dagg):
        h, weight = ctx.saved_tensors
        g = ctx.g
        N, d_in = h.shape
        R2 = 2 * g.num_rels
        type_rowptr, type_src, type_dst = _block_index(g)
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        nb = _lib.load().regcn_block_aggregate_bwd_workspace_bytes(R2, d_in, ctx.d_out, ctx.nb)
        ws = _ws(h.device, nb)
        call("regcn_block_aggregate_bwd", ptr(h), ptr(dagg.contiguous()), ptr(weight), ptr(g.rowptr), ptr(g.src_sorted),
             ptr(g.etype_sorted), ptr(g.norm), ptr(type_rowptr), ptr(type_src), ptr(type_dst), N, R2, d_in, ctx.d_out,
             ctx.nb, ptr(dh), ptr(dW), ptr(ws), nb)
        return dh, dW, None, None, None


class _RReluDrop(torch.autograd.Function):
    """out = dropout_p(rrelu(x))  -- the activation of a layer without self-loop (rgcn/layers.py:84-87)."""

    @staticmethod
    def forward(ctx, x, p):
        out, _, _ = ops.union_combine(x.contiguous(), None, None, act=1)
        if p > 0:
            call("regcn_dropout", ptr(out), out.numel(), float(p), _next_seed())
        ctx.p = float(p)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        N, d = out.shape
        dx = torch.empty_like(out)
        call("regcn_union_combine_bwd", ptr(out), ptr(dout.contiguous()), None, N, d, ctx.p, ptr(dx), None)
        return dx, None


def static_angle_terms(static_emb, hist, layer_norm, angle, discount, weight):
    """src/rrgcn.py:225-247 forward: (loss (1,), cosines) -- one row kernel per history step + one fixed-order sum."""
    import math
    N, d = static_emb.shape
    L = len(hist)
    dev = static_emb.device
    terms = torch.empty((max(L, 1) * N, 1), device=dev, dtype=F32)
    coss = []
    for t, e in enumerate(hist):
        step = (angle * math.pi / 180) * ((t + 1) if discount == 1 else 1)
        coss.append(math.cos(step))
        call("regcn_static_angle_fwd", ptr(static_emb), ptr(e.contiguous()), N, d, coss[-1], float(weight),
             int(bool(layer_norm)), terms[t * N:(t + 1) * N].data_ptr())
    if L == 0:
        return torch.zeros(1, device=dev), coss
    return _col_sum(terms), coss


class _StaticAngle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, static_emb, layer_norm, angle, discount, weight, *hist):
        static_emb = static_emb.contiguous()
        hist = [e.contiguous() for e in hist]
        loss, coss = static_angle_terms(static_emb, hist, layer_norm, angle, discount, weight)
        ctx.save_for_backward(static_emb, *hist)
        ctx.cfg = (bool(layer_norm), float(weight), coss)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        static_emb, *hist = ctx.saved_tensors
        layer_norm, weight, coss = ctx.cfg
        N, d = static_emb.shape
        dS = torch.zeros_like(static_emb)
        dEs = []
        g = gloss.contiguous().view(-1)
        for t, e in enumerate(hist):
            dE = torch.empty_like(e)
            call("regcn_static_angle_bwd", ptr(static_emb), ptr(e), N, d, coss[t], weight, int(layer_norm), ptr(g),
                 ptr(dS), 1, ptr(dE))
            dEs.append(dE)
        return (dS, None, None, None, None, *dEs)


def static_embedding(model, static_graph):
    """src/rrgcn.py:146-152 (hyperbolic_model.py:762-770) with the tape on: block layer over cat(dynamic_emb, words_emb),
    entity rows, F.normalize."""
    layer = model.statci_rgcn_layer if hasattr(model, "statci_rgcn_layer") else model.static_rgcn_layer
    x = torch.cat((model.dynamic_emb, model.words_emb), dim=0)
    agg = _BlockAggregate.apply(x, layer.weight, static_graph, layer.num_bases, layer.out_feat)
    out = _RReluDrop.apply(agg, 0.0)              # RGCNLayer drops only the self-loop message (:52-53); none here
    s_emb = out[:model.num_ents]
    return normalize(s_emb) if model.layer_norm else s_emb.contiguous()


rel_mean_pool = _RelMeanPool.apply
union_aggregate = _UnionAggregate.apply
union_combine = _UnionCombine.apply
union_sum = _UnionSum.apply
rrelu_drop = _RReluDrop.apply
time_gate = _TimeGate.apply
gru_gate = _GRUGate.apply
normalize = _Normalize.apply
tanh = _Tanh.apply


# ------------------------------------------------------------------------------------------------ model-level glue
def regcn_evolve(model, g_list, static_graph=None):
    """RecurrentRGCN.forward with the tape on (src/rrgcn.py:142-180; uvrgcn, self_loop).  --skip-connect changes
    nothing here: RGCNCell calls every layer with prev_h=[] (src/rrgcn.py:37-38), so the gate of rgcn/layers.py:234-245
    is never taken and its weights never receive a gradient.  Returns (history_embs, h_0, static_emb)."""
    if not model.rgcn.self_loop or model.encoder_name != "uvrgcn":
        raise NotImplementedError("regcn_b200.train: uvrgcn + self_loop only")
    cell = model.relation_cell_1
    static_emb = None
    if model.use_static:
        static_emb = static_embedding(model, static_graph)
        h = static_emb
    else:
        h = normalize(model.dynamic_emb) if model.layer_norm else model.dynamic_emb
    h0 = None
    hist = []
    loop_cats = [loop_cat(layer) for layer in model.rgcn.layers]
    for i, g in enumerate(g_list):
        x_mean = rel_mean_pool(h, g)
        x_cat = torch.cat((model.emb_rel, x_mean), dim=1)
        gi = linear(x_cat, cell.weight_ih, cell.bias_ih)
        hprev = model.emb_rel if i == 0 else h0
        gh = linear(hprev, cell.weight_hh, cell.bias_hh)
        h0 = gru_gate(gi, gh, hprev, model.layer_norm)
        cur = h
        for layer, lw in zip(model.rgcn.layers, loop_cats):
            p = float(layer.dropout.p) if (layer.dropout is not None and model.training) else 0.0
            agg = union_aggregate(cur, h0, g)
            P = linear(agg, layer.weight_neighbor, None, True)
            L = linear(cur, lw, None, True)
            cur = union_combine(P, L, g, p)
        G = linear(h, model.time_gate_weight, None, True)
        h = time_gate(G, model.time_gate_bias, cur, h, model.layer_norm)
        hist.append(h)
    return hist, h0, static_emb


def regcn_get_loss(model, glist, triples, static_graph=None):
    """src/rrgcn.py:197-223 with gradients: (loss_ent, loss_rel, loss_static), each of shape (1,)."""
    _lib.require_device()
    begin_step()
    if ops.gemm_impl() != "tc":
        raise RuntimeError("regcn_b200.train needs the tensor-core GEMM (REGCN_GEMM=tc)")
    dev = model.dynamic_emb.device
    triples = torch.as_tensor(triples).to(dev)
    inverse = triples.flip(1)
    inverse[:, 1] = inverse[:, 1] + model.num_rels
    all_triples = torch.cat([triples, inverse]).contiguous()
    hist, r_emb, static_emb = regcn_evolve(model, glist, static_graph)
    pre = normalize(hist[-1]) if model.layer_norm else hist[-1]
    e_all = tanh(pre)
    loss_ent = torch.zeros(1, device=dev)
    loss_rel = torch.zeros(1, device=dev)
    loss_static = torch.zeros(1, device=dev)
    if model.entity_prediction:
        q = conv_tower(model.decoder_ob, e_all, r_emb, all_triples, 0, 1)
        loss_ent = score_ce(q, e_all, all_triples, 2)
    if model.relation_prediction:
        q = conv_tower(model.rdecoder, e_all, e_all, all_triples, 0, 2)
        loss_rel = score_ce(q, r_emb, all_triples, 1)
    if model.use_static and model.discount in (0, 1):
        loss_static = _StaticAngle.apply(static_emb, model.layer_norm, model.angle, model.discount, model.weight, *hist)
    return loss_ent, loss_rel, loss_static
