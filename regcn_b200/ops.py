"""Thin tensor-level wrappers over the C ABI (one per kernel entry point).

Each wrapper allocates its outputs with torch (device memory ownership stays with the caller's
framework), validates dtypes, and launches on torch's current stream.  No arithmetic happens here.
"""
import torch

from . import _lib
from ._lib import call, ptr

F32 = torch.float32
I32 = torch.int32
I64 = torch.int64


def _f32(t, name):
    if t.dtype != F32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _rowmajor(t, name):
    """fp32 2-D tensor with unit column stride (row slices / column blocks of a bigger matrix are fine)."""
    if t.dtype != F32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D")
    if not t.is_cuda:
        raise RuntimeError("regcn_b200: kernels take CUDA tensors only (no CPU fallback)")
    if t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


# --------------------------------------------------------------------------- dense contractions
import os

# ONE dense-contraction backend on the product path: the tcgen05 kernel (csrc/gemm_tc.cu).  "tc" = 3xTF32
# error-compensated (fp32 parity, the default), "tc1" = single TF32 pass (reported separately, never the parity path).
# The fp32 CUDA-core kernel (csrc/gemm_simt.cu) is NOT selectable here: it is the yardstick tests/ compare against
# through gemm_f32_yardstick() below.
_GEMM_IMPL = {"impl": "tc"}
_IMPLS = {"tc": "regcn::tc::gemm_tf32_kernel (tcgen05 kind::tf32, 3xTF32 error-compensated)",
          "tc1": "regcn::tc::gemm_tf32_kernel (tcgen05 kind::tf32, single pass)"}


def set_gemm_impl(name):
    """tc: tcgen05 3xTF32 (fp32 parity); tc1: tcgen05 plain TF32 (one pass)."""
    if name not in _IMPLS:
        raise ValueError(f"unknown gemm implementation {name!r} (the product path has one backend: tc | tc1)")
    _GEMM_IMPL["impl"] = name


def gemm_impl():
    return _GEMM_IMPL["impl"]


# all-entity scoring precision: "fp32" = 3xTF32 error-compensated (the parity mode), "bf16" = tcgen05 kind::f16 on
# bf16-rounded operands with fp32 accumulation (reported separately: ranks may move where scores are within ~1e-2)
_SCORE_DTYPE = {"dtype": os.environ.get("REGCN_SCORE_DTYPE", "fp32")}


def set_score_dtype(name):
    if name not in ("fp32", "bf16"):
        raise ValueError(f"unknown scoring dtype {name!r}")
    _SCORE_DTYPE["dtype"] = name


def score_dtype():
    return _SCORE_DTYPE["dtype"]


def to_bf16(x):
    """(M, K) fp32 -> bf16 rows (K % 8 == 0), round to nearest even."""
    M, K = x.shape
    if K % 8:
        raise ValueError("to_bf16: K must be a multiple of 8")
    out = torch.empty((M, K), device=x.device, dtype=torch.bfloat16)
    call("regcn_to_bf16", ptr(x.contiguous()), ptr(out), M * K)
    return out


def gemm_kernel_name():
    return _IMPLS[_GEMM_IMPL["impl"]]


def cached(owner, key, fn):
    """Derived static operand of a parameter (transpose / slice / tf32 split), kept on the tensor object itself and
    recomputed when its version or storage changes.  `owner` must be a long-lived tensor object (the Parameter)."""
    store = owner.__dict__.setdefault("_regcn_cache", {})
    stamp = (owner._version, owner.data_ptr())
    hit = store.get(key)
    if hit is None or hit[0] != stamp:
        with torch.no_grad():
            store[key] = (stamp, fn())
    return store[key][1]


def split_tf32(x):
    """(hi, lo) TF32 split of a contiguous fp32 tensor (numel % 4 == 0)."""
    hi = torch.empty_like(x)
    lo = torch.empty_like(x)
    call("regcn_split_tf32", ptr(x), ptr(hi), ptr(lo), x.numel())
    return hi, lo


def _tc_operand(m, need_lo):
    m = m.contiguous()
    if m.data_ptr() % 16:
        m = m.clone()
    if not need_lo:
        return m, None
    return split_tf32(m)


def gemm(a, b, trans_b=False, bias=None, out=None, accumulate=False, split_k=1, b_key=None, split_a_on_chip=False):
    """C[M,N] (+)= A[M,K] @ (B[K,N] | B[N,K]^T) (+ bias) on the tcgen05 kernel.  torch.mm / F.linear call sites of the path.

    split_a_on_chip: A stays one fp32 matrix in HBM; TMA brings its k-blocks into the operand ring and two converter warps
    split them to TF32 (hi, lo) there (regcn_gemm_tf32_a32) -- the same operand values, half the bytes of A.

    b_key=(owner_tensor, name): B is a static weight; its prepared form (K-major transpose, TF32 split) is cached on
    `owner_tensor` under `name`.  K must be a multiple of 4 (16-byte operand rows for TMA): every contraction of the
    path has K = h_dim, 2 h_dim or 50 h_dim; anything else raises -- there is no second backend to fall to."""
    impl = _GEMM_IMPL["impl"]
    K = (a[0] if isinstance(a, tuple) else a).shape[1]
    if K % 4:
        raise ValueError(f"gemm: reduction length K={K} must be a multiple of 4 (TMA needs 16-byte operand rows); pad the "
                         "operands with zero columns")
    return _gemm_tc(a, b, trans_b, bias, out, accumulate, split_k, b_key, 3 if impl == "tc" else 1, split_a_on_chip)


def gemm_f32_yardstick(a, b, trans_b=False, bias=None, out=None, accumulate=False, split_k=1):
    """Plain fp32 CUDA-core GEMM (csrc/gemm_simt.cu, regcn_gemm_f32).  TEST / MEASUREMENT YARDSTICK ONLY: nothing in
    regcn_b200/ calls it; tests compare the tensor-core path against it and against fp64."""
    a = _rowmajor(a, "a")
    b = _rowmajor(b, "b")
    M, K = a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    if (b.shape[1] if trans_b else b.shape[0]) != K:
        raise ValueError(f"gemm: inner dims differ {tuple(a.shape)} x {tuple(b.shape)} trans_b={trans_b}")
    if out is None:
        if accumulate:
            raise ValueError("gemm: accumulate needs out")
        out = torch.empty((M, N), device=a.device, dtype=F32)
    elif not out.is_cuda or out.dtype != F32 or out.stride(1) != 1:
        raise ValueError("gemm: out must be a CUDA float32 matrix with unit column stride")
    ldc = out.stride(0)
    ws = None
    ws_bytes = 0
    if split_k > 1:
        ws_bytes = _lib.load().regcn_gemm_f32_workspace_bytes(M, N, split_k)
        ws = torch.empty(ws_bytes // 4, device=a.device, dtype=F32)
    call("regcn_gemm_f32", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), int(trans_b), out.data_ptr(), ldc, M, N, K,
         ptr(bias), int(accumulate), split_k, ptr(ws), ws_bytes)
    return out


def _gemm_tc(a, b, trans_b, bias, out, accumulate, split_k, b_key, passes, a_on_chip=False):
    a_pre = a if isinstance(a, tuple) else None      # (hi, lo) produced directly by the upstream kernel
    if a_pre is not None:
        a = a_pre[0]
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("regcn_b200: kernels take CUDA tensors only (no CPU fallback)")
    if a.dtype != F32 or b.dtype != F32:
        raise TypeError("gemm: operands must be float32")
    M, K = a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    if (b.shape[1] if trans_b else b.shape[0]) != K:
        raise ValueError(f"gemm: inner dims differ {tuple(a.shape)} x {tuple(b.shape)} trans_b={trans_b}")
    need_lo = passes == 3

    def prep_b():
        bd = b.detach()
        return _tc_operand(bd if trans_b else bd.t(), need_lo)

    if b_key is not None:
        owner, name = b_key
        b_hi, b_lo = cached(owner, ("tcB", name, trans_b, passes, tuple(b.shape), b.storage_offset(), tuple(b.stride())),
                            prep_b)
    else:
        b_hi, b_lo = prep_b()
    a_on_chip = a_on_chip and a_pre is None
    if a_on_chip:
        a = a.detach()
        if a.stride(1) != 1 or a.stride(0) % 4 or a.data_ptr() % 16:
            a = a.contiguous()
    else:
        a_hi, a_lo = a_pre if a_pre is not None else _tc_operand(a.detach(), need_lo)
    if out is None:
        if accumulate:
            raise ValueError("gemm: accumulate needs out")
        if N % 4:
            out = torch.empty((M, (N + 3) // 4 * 4), device=a.device, dtype=F32)[:, :N]   # 16-byte aligned rows
        else:
            out = torch.empty((M, N), device=a.device, dtype=F32)
    elif not out.is_cuda or out.dtype != F32 or out.stride(1) != 1:
        raise ValueError("gemm: out must be a CUDA float32 matrix with unit column stride")
    ws = None
    ws_bytes = 0
    if split_k > 1:
        ws_bytes = _lib.load().regcn_gemm_tf32_workspace_bytes(M, N, split_k)
        ws = torch.empty(ws_bytes // 4, device=a.device, dtype=F32)
    if a_on_chip:
        call("regcn_gemm_tf32_a32", a.data_ptr(), a.stride(0), K, None, None, 4, 0, None, ptr(b_hi), ptr(b_lo), K,
             out.data_ptr(), out.stride(0), M, N, ptr(bias), int(accumulate), passes, split_k, ptr(ws), ws_bytes, None, 0)
        return out
    call("regcn_gemm_tf32", ptr(a_hi), ptr(a_lo), K, ptr(b_hi), ptr(b_lo), K, out.data_ptr(), out.stride(0), M, N, K,
         ptr(bias), int(accumulate), passes, split_k, ptr(ws), ws_bytes)
    return out


# --------------------------------------------------------------------------- edge path
def rel_mean_pool(h, g, nsplit=None):
    h = _f32(h, "h")
    R, d = g.num_rels, h.shape[1]
    out = torch.empty((2 * R, d), device=h.device, dtype=F32)
    if nsplit is None:
        nsplit = max(1, min(64, g.n_rel_ents // (R * 512)))
    partial = torch.empty((R * nsplit, d), device=h.device, dtype=F32) if nsplit > 1 else None
    call("regcn_rel_mean_pool", ptr(h), ptr(g.rel_rowptr), ptr(g.rel_ents), R, d, nsplit, ptr(out), ptr(partial))
    return out


def union_aggregate(h, rel, g, radius=None, gamma=0.0, out=None):
    h = _f32(h, "h")
    rel = _f32(rel, "rel")
    N, d = h.shape
    if out is None:
        out = torch.empty((N, d), device=h.device, dtype=F32)
    partial = torch.empty((g.n_split_chunks, d), device=h.device, dtype=F32) if g.n_split_chunks > 0 else None
    call("regcn_union_aggregate", ptr(h), ptr(rel), ptr(g.rowptr), ptr(g.src_sorted), ptr(g.etype_sorted), ptr(g.norm),
         ptr(g.vptr), ptr(g.sptr), ptr(g.vrow_row), g.n_vrows, g.n_split_chunks, ptr(radius), float(gamma), N, d,
         ptr(out), ptr(partial))
    return out


def block_aggregate(h, weight, g, num_bases, d_out, radius=None, gamma=0.0):
    """K6; radius (N,) + gamma: messages weighted by exp(-gamma |radius[src] - radius[dst]|) (HyperbolicRGCNLayer)."""
    h = _f32(h, "h")
    weight = _f32(weight, "weight")
    N, d_in = h.shape
    out = torch.empty((N, d_out), device=h.device, dtype=F32)
    if radius is not None:
        call("regcn_block_aggregate_radius", ptr(h), ptr(weight), ptr(_f32(radius, "radius")), float(gamma), ptr(g.rowptr),
             ptr(g.src_sorted), ptr(g.etype_sorted), ptr(g.norm), N, d_in, d_out, num_bases, ptr(out))
    else:
        call("regcn_block_aggregate", ptr(h), ptr(weight), ptr(g.rowptr), ptr(g.src_sorted), ptr(g.etype_sorted),
             ptr(g.norm), N, d_in, d_out, num_bases, ptr(out))
    return out


def lorentz_aggregate(ht, weight, rel, g, num_bases, c):
    ht = _f32(ht, "ht")
    N, d = ht.shape
    out = torch.empty((N, d), device=ht.device, dtype=F32)
    partial = torch.empty(g.n_split_chunks * (d + 1), device=ht.device, dtype=F32) if g.n_split_chunks > 0 else None
    call("regcn_lorentz_aggregate", ptr(ht), ptr(_f32(weight, "weight")), ptr(rel), ptr(g.rowptr), ptr(g.src_sorted),
         ptr(g.etype_sorted), ptr(g.norm), ptr(g.vptr), ptr(g.sptr), ptr(g.vrow_row), g.n_vrows, g.n_split_chunks, N, d,
         num_bases, float(c), ptr(out), ptr(partial))
    return out


# --------------------------------------------------------------------------- row maps
ROW_NORMALIZE, ROW_TANH, ROW_LEAKY_TANH_LOG0, ROW_LOG0, ROW_EXP0, ROW_PROJECT, ROW_TANGENT_NORMALIZE, ROW_IDENTITY, \
    ROW_RRELU_EXP0, ROW_NORMALIZE_TANH = range(10)


def row_sumsq(x):
    """|x_n|^2 per row (the e_sumsq operand of the norm/dot score form)."""
    x = _f32(x, "x")
    M, d = x.shape
    ss = torch.empty(M, device=x.device, dtype=F32)
    call("regcn_row_map", ptr(x), None, M, d, ROW_IDENTITY, 1.0, ptr(ss))
    return ss


def row_map(x, mode, c=1.0, want_sumsq=False, out=None, split=False):
    """split=True: the TF32 (hi, lo) pair of the result is produced in the same pass and rides on the returned tensor as
    `_regcn_split` (consumed by fused_rank_counts instead of a separate split kernel)."""
    x = _f32(x, "x")
    M, d = x.shape
    if out is None:
        out = torch.empty_like(x)
    if split and not want_sumsq:
        hi, lo = torch.empty_like(out), torch.empty_like(out)
        call("regcn_row_map_split", ptr(x), ptr(out), ptr(hi), ptr(lo), M, d, mode, float(c))
        out._regcn_split = (hi, lo)
        return out
    ss = torch.empty(M, device=x.device, dtype=F32) if want_sumsq else None
    call("regcn_row_map", ptr(x), ptr(out), M, d, mode, float(c), ptr(ss))
    return (out, ss) if want_sumsq else out


def gru_gate(gi, gh, hprev, normalize):
    M, d = hprev.shape
    out = torch.empty((M, d), device=hprev.device, dtype=F32)
    call("regcn_gru_gate", ptr(gi), ptr(gh), ptr(_f32(hprev, "hprev")), ptr(out), M, d, int(normalize))
    return out


def union_combine(P, L, indeg, act=1, hyper=False, c=1.0, skip=None, skip_bias=None, prev=None,
                  want_tangent=False, want_radius=False):
    N, d = P.shape
    out = torch.empty((N, d), device=P.device, dtype=F32)
    ht = torch.empty((N, d), device=P.device, dtype=F32) if want_tangent else None
    rad = torch.empty(N, device=P.device, dtype=F32) if want_radius else None
    call("regcn_union_combine", ptr(P), ptr(L), ptr(indeg), ptr(skip), ptr(skip_bias), ptr(prev), N, d, act,
         int(hyper), float(c), ptr(out), ptr(ht), ptr(rad))
    return out, ht, rad


def time_gate(G, bias, cur, h, normalize_cur):
    N, d = h.shape
    out = torch.empty((N, d), device=h.device, dtype=F32)
    call("regcn_time_gate", ptr(G), ptr(_f32(bias, "bias")), ptr(cur), ptr(_f32(h, "h")), ptr(out), N, d,
         int(normalize_cur))
    return out


def hyp_init(emb, radius_static, normalize, on_manifold, c, rmin, rmax):
    emb = _f32(emb, "emb")
    N, d = emb.shape
    out = torch.empty((N, d), device=emb.device, dtype=F32)
    call("regcn_hyp_init", ptr(emb), ptr(radius_static), N, d, int(normalize), int(on_manifold), float(c), float(rmin),
         float(rmax), ptr(out))
    return out


def hyp_tangent(h, c, want_clamped=True, want_radius=True):
    h = _f32(h, "h")
    N, d = h.shape
    ht = torch.empty((N, d), device=h.device, dtype=F32)
    pt = torch.empty((N, d), device=h.device, dtype=F32) if want_clamped else None
    rad = torch.empty(N, device=h.device, dtype=F32) if want_radius else None
    call("regcn_hyp_tangent", ptr(h), N, d, float(c), ptr(ht), ptr(pt), ptr(rad))
    return ht, pt, rad


def hyp_time_gate(h2, pt, G, bias, radius_static, radius_w, radius_b, layer_norm, residual, c, rmin, rmax, beta, eps_r):
    N, d = h2.shape
    out = torch.empty((N, d), device=h2.device, dtype=F32)
    call("regcn_hyp_time_gate", ptr(h2), ptr(pt), ptr(G), ptr(_f32(bias, "bias")), ptr(_f32(radius_static, "rs")),
         ptr(radius_w), float(radius_b), N, d, int(layer_norm), int(residual), float(c), float(rmin), float(rmax),
         float(beta), float(eps_r), ptr(out))
    return out


# --------------------------------------------------------------------------- decoders
def convtranse_features(ent, second, triples, col0, col1, bn0, conv_w, conv_b, bn1, split=False):
    """bn0 / bn1 = (scale, shift) folded eval-mode BatchNorm.  split=True returns the TF32 (hi, lo) pair the FC GEMM
    consumes instead of the raw feature matrix (no separate conversion pass)."""
    B = triples.shape[0]
    d = ent.shape[1]
    C, _, ksz = conv_w.shape
    dev = ent.device
    if split:
        hi = torch.empty((B, C * d), device=dev, dtype=F32)
        lo = torch.empty((B, C * d), device=dev, dtype=F32)
        call("regcn_convtranse_features", ptr(ent), ptr(second), ptr(triples), col0, col1, B, d, C, ksz, ptr(bn0[0]),
             ptr(bn0[1]), ptr(conv_w.contiguous()), ptr(conv_b), ptr(bn1[0]), ptr(bn1[1]), None, ptr(hi), ptr(lo))
        return hi, lo
    out = torch.empty((B, C * d), device=dev, dtype=F32)
    call("regcn_convtranse_features", ptr(ent), ptr(second), ptr(triples), col0, col1, B, d, C, ksz, ptr(bn0[0]),
         ptr(bn0[1]), ptr(conv_w.contiguous()), ptr(conv_b), ptr(bn1[0]), ptr(bn1[1]), ptr(out), None, None)
    return out


def convtrans_fc_ok(d, conv_w, n_out):
    """Shapes regcn_convtrans_fc takes (kernel size 3, d % 4 == 0, at most 64 channels); REGCN_FUSED_TOWER=0 turns it off."""
    import os
    C, _, ksz = conv_w.shape
    return ksz == 3 and d % 4 == 0 and C <= 64 and n_out % 4 == 0 and os.environ.get("REGCN_FUSED_TOWER", "1") != "0"


def convtrans_fc_weight(fc_w, C, d):
    """fc.weight (n_out, C d) in the reduction order of regcn_convtrans_fc (blocks of 16 positions outermost), TF32 split;
    cached on the Parameter until it changes."""
    def build():
        w = fc_w.detach().contiguous()
        n_out = w.shape[0]
        kp = 16 * C * ((d + 15) // 16)
        hi = torch.empty((n_out, kp), device=w.device, dtype=F32)
        lo = torch.empty((n_out, kp), device=w.device, dtype=F32)
        call("regcn_convtrans_fc_pack_weight", ptr(w), n_out, C, d, ptr(hi), ptr(lo))
        return hi, lo
    return cached(fc_w, ("convfc", C, d, tuple(fc_w.shape)), build)


def convtrans_fc(ent, second, triples, col0, col1, bn0, conv_w, conv_b, bn1, fc_w, fc_b, batch_total=None, bn2=None,
                 relu=False):
    """bn0 -> conv1d -> bn1 -> relu -> fc in one tcgen05 GEMM whose A operand (the feature map) is computed inside the
    operand ring (regcn_convtrans_fc).  fc_w: the fc.weight Parameter itself (its TF32 split is cached on it, shared with
    ops.gemm(..., b_key=(fc_w, "w"))).  bn2 = (scale, shift) / relu: the tower's tail folded into the split-K reduction.
    Returns the (B, n_out) activations."""
    B = triples.shape[0]
    d = ent.shape[1]
    C, _, ksz = conv_w.shape
    n_out = fc_w.shape[0]
    w_hi, w_lo = convtrans_fc_weight(fc_w, C, d)
    total = int(batch_total) if batch_total else B
    ws_bytes = _lib.load().regcn_convtrans_fc_workspace_bytes(total, n_out)
    ws = torch.empty(ws_bytes // 4 + 1, device=ent.device, dtype=F32)
    out = torch.empty((B, n_out), device=ent.device, dtype=F32)
    call("regcn_convtrans_fc", ptr(ent), ptr(second), ptr(triples), col0, col1, B, total, d, C, ksz, ptr(bn0[0]), ptr(bn0[1]),
         ptr(conv_w.contiguous()), ptr(conv_b), ptr(bn1[0]), ptr(bn1[1]), ptr(w_hi), ptr(w_lo), w_hi.shape[1], n_out, ptr(fc_b),
         ptr(bn2[0]) if bn2 is not None else None, ptr(bn2[1]) if bn2 is not None else None, int(relu), ptr(out), n_out,
         None, None, ptr(ws), ws_bytes)
    return out


def affine_relu_(x, scale, shift, relu=True):
    M, d = x.shape
    call("regcn_affine_relu", ptr(x), ptr(scale), ptr(shift), M, d, int(relu))
    return x


def gather_log0(E, triples, col, project, c):
    B = triples.shape[0]
    d = E.shape[1]
    out = torch.empty((B, d), device=E.device, dtype=F32)
    call("regcn_gather_log0", ptr(E), ptr(triples), col, B, d, int(project), float(c), ptr(out))
    return out


def hyp_query(s_tan, ang, trans, E, triples, kind, c):
    B, d = s_tan.shape
    Q = torch.empty((B, d), device=s_tan.device, dtype=F32)
    qss = torch.empty(B, device=s_tan.device, dtype=F32)
    call("regcn_hyp_query", ptr(s_tan), ptr(ang), ptr(trans), ptr(E), ptr(triples), B, d, kind, float(c), ptr(Q),
         ptr(qss))
    return Q, qss


def rel_curvature(raw, triples, num_relations, c, cmax):
    """Per-query curvature of the relation-specific-curvature decoders (hyperbolic_decoder.py:66-86,1020-1026)."""
    B = triples.shape[0]
    out = torch.empty(B, device=triples.device, dtype=F32)
    call("regcn_rel_curvature", ptr(raw.detach().contiguous()), ptr(triples), B, int(num_relations), float(c),
         float(cmax) if cmax is not None else 0.0, ptr(out))
    return out


def hyp_score_epilogue_(S, q_sumsq, e_sumsq, bias, qbias, c, scale_margin, row_c=None):
    B, N = S.shape
    step = 65535
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        call("regcn_hyp_score_epilogue", S[b0:b1].data_ptr(), S.stride(0), b1 - b0, N, q_sumsq[b0:b1].data_ptr(),
             ptr(e_sumsq), ptr(bias), None if qbias is None else qbias[b0:b1].data_ptr(), float(c), ptr(scale_margin),
             None if row_c is None else row_c[b0:b1].data_ptr())
    return S


# --------------------------------------------------------------------------- ranking
def rank_dense(score, triples, target_col, filt_ptr=None, filt_idx=None, col_offset=0, target_score=None,
               filt_end=None):
    """Counts for a (B, N_shard) dense score block.  Returns (raw_count, filt_count, target_score) int32/int32/f32."""
    B, N = score.shape
    dev = score.device
    if not score.is_cuda or score.dtype != F32 or score.stride(1) != 1:
        raise ValueError("rank_dense: score must be a CUDA float32 matrix with unit column stride")
    if target_score is None:
        target_score = torch.zeros(B, device=dev, dtype=F32)
        call("regcn_gather_target_score", score.data_ptr(), score.stride(0), B, N, ptr(triples), target_col, col_offset,
             ptr(target_score))
    raw = torch.empty(B, device=dev, dtype=I32)
    filt = torch.empty(B, device=dev, dtype=I32) if filt_ptr is not None else None
    call("regcn_rank_count", score.data_ptr(), score.stride(0), B, N, ptr(triples), target_col, ptr(filt_ptr),
         ptr(filt_idx), col_offset, ptr(target_score), ptr(raw), ptr(filt), ptr(filt_end))
    return raw, filt, target_score


def counts_to_ranks(raw, filt):
    B = raw.shape[0]
    rank = torch.empty(B, device=raw.device, dtype=I64)
    frank = torch.empty(B, device=raw.device, dtype=I64)
    call("regcn_counts_to_ranks", ptr(raw), ptr(filt), B, ptr(rank), ptr(frank))
    return rank, frank


def apply_filter_(score, triples, target_col, filt_ptr, filt_idx, col_offset=0, filt_end=None):
    B, N = score.shape
    call("regcn_apply_filter", score.data_ptr(), score.stride(0), B, N, ptr(triples), target_col, ptr(filt_ptr),
         ptr(filt_idx), col_offset, ptr(filt_end))
    return score


# --------------------------------------------------------------------------- fused scoring + rank (no score matrix)
def fused_rank_counts(q, cand, target, filt_ptr, filt_idx, pair_a, pair_e, hyp=None, col_bias=None, shard=None,
                      cand_split=None, filt_end=None):
    """Raw / filtered 'beats the target' counts of every query against the candidate rows [lo,hi) of `cand`, computed by
    the scoring GEMM's counting epilogue (K11/K13 fused with K14); the (B,N) score matrix is never written.

    q (B,d), cand (N,d) fp32; target (B,) int32 global candidate ids; filt_ptr/filt_idx the filter CSR;
    pair_a/pair_e int32 pair lists: first the B (query, target) pairs, then one pair per filter-CSR entry;
    hyp = (c, q_sumsq, e_sumsq, scale_margin[, row_c]) selects the hyperbolic score (row_c (B,): per-query curvature,
    true-distance artanh branch); col_bias (N,) optional candidate bias.
    Returns (raw_count, filt_count, target_score)."""
    if _GEMM_IMPL["impl"] not in ("tc", "tc1"):
        raise RuntimeError("fused_rank_counts needs the tensor-core GEMM (REGCN_GEMM=tc)")
    passes = 3 if _GEMM_IMPL["impl"] == "tc" else 1
    B, K = q.shape
    N = cand.shape[0]
    dev = q.device
    P = pair_a.shape[0]
    bf16 = _SCORE_DTYPE["dtype"] == "bf16"
    if bf16:
        # bf16 mode: one kind::f16 pass on bf16-rounded operands; the pair pass uses the same operands and the same
        # MMA arithmetic, so target / filter-entry scores stay bit-identical to the counted scores
        passes = 0
        q_hi, q_lo = to_bf16(q.detach()), None
        e_hi, e_lo = (cand_split if cand_split is not None else (to_bf16(cand.detach()), None))
        a_hi = torch.empty((P, K), device=dev, dtype=torch.bfloat16)
        b_hi = torch.empty((P, K), device=dev, dtype=torch.bfloat16)
        a_lo = b_lo = None
        call("regcn_gather_rows2", ptr(q_hi), None, ptr(pair_a), P, K // 2, ptr(a_hi), None)     # a bf16 row = K/2 words
        call("regcn_gather_rows2", ptr(e_hi), None, ptr(pair_e), P, K // 2, ptr(b_hi), None)
    else:
        q_hi, q_lo = _tc_operand(q.detach(), passes == 3)
        if cand_split is None and passes == 3:
            cand_split = getattr(cand, "_regcn_split", None)       # produced with the table (row_map(split=True))
        e_hi, e_lo = cand_split if cand_split is not None else _tc_operand(cand.detach(), passes == 3)
        a_hi = torch.empty((P, K), device=dev, dtype=F32)
        b_hi = torch.empty((P, K), device=dev, dtype=F32)
        a_lo = torch.empty((P, K), device=dev, dtype=F32) if passes == 3 else None
        b_lo = torch.empty((P, K), device=dev, dtype=F32) if passes == 3 else None
        call("regcn_gather_rows2", ptr(q_hi), ptr(q_lo), ptr(pair_a), P, K, ptr(a_hi), ptr(a_lo))
        call("regcn_gather_rows2", ptr(e_hi), ptr(e_lo), ptr(pair_e), P, K, ptr(b_hi), ptr(b_lo))
    c, x2, y2, sm = (hyp[:4] if hyp is not None else (1.0, None, None, None))
    row_c = hyp[4] if hyp is not None and len(hyp) > 4 else None
    x2p = y2p = bp = rcp = None
    if hyp is not None or col_bias is not None:
        x2p = torch.empty(P, device=dev, dtype=F32) if hyp is not None else None
        y2p = torch.empty(P, device=dev, dtype=F32) if hyp is not None else None
        bp = torch.empty(P, device=dev, dtype=F32) if col_bias is not None else None
        call("regcn_gather_scalars", ptr(x2), ptr(y2), ptr(col_bias), ptr(pair_a), ptr(pair_e), P, ptr(x2p), ptr(y2p),
             ptr(bp))
        if row_c is not None:
            rcp = torch.empty(P, device=dev, dtype=F32)
            call("regcn_gather_scalars", ptr(row_c), None, None, ptr(pair_a), ptr(pair_e), P, ptr(rcp), None, None)
    ps = torch.empty(P, device=dev, dtype=F32)
    call("regcn_pair_scores_tf32", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), P, K, int(hyp is not None), ptr(x2p),
         ptr(y2p), ptr(bp), float(c), ptr(sm), ptr(rcp), ptr(ps), passes)
    lo, hi = shard if shard is not None else (0, N)
    raw = torch.zeros(B, device=dev, dtype=I32)
    if hi > lo:
        call("regcn_score_count_tf32", ptr(q_hi), ptr(q_lo), e_hi[lo:hi].data_ptr(),
             e_lo[lo:hi].data_ptr() if e_lo is not None else None, B, hi - lo, K, ptr(ps), ptr(target), ptr(raw), lo,
             int(hyp is not None), ptr(x2), y2[lo:hi].data_ptr() if y2 is not None else None,
             col_bias[lo:hi].data_ptr() if col_bias is not None else None, float(c), ptr(sm), ptr(row_c), passes)
    filt = torch.empty(B, device=dev, dtype=I32)
    call("regcn_filter_correct", B, ptr(filt_ptr), ptr(filt_idx), ptr(target), ptr(ps), ptr(raw), lo, hi, ptr(filt),
         ptr(filt_end))
    return raw, filt, ps[:B]


# --------------------------------------------------------------------------- loss heads (forward)
def fused_ce(q, cand, target, hyp=None, col_bias=None):
    """Cross entropy of every query row against ALL candidates without the (B,N) logits: the scoring GEMM runs with a
    streaming log-sum-exp epilogue, the target logit comes from the pair-score pass (same arithmetic).
    Returns (ce (B,), loss (1,) = mean(ce)).  Arguments as fused_rank_counts."""
    if _GEMM_IMPL["impl"] not in ("tc", "tc1"):
        raise RuntimeError("fused_ce needs the tensor-core GEMM (REGCN_GEMM=tc)")
    passes = 3 if _GEMM_IMPL["impl"] == "tc" else 1
    B, K = q.shape
    N = cand.shape[0]
    dev = q.device
    target = target.to(torch.int32).contiguous()
    q_hi, q_lo = _tc_operand(q.detach(), passes == 3)
    e_hi, e_lo = _tc_operand(cand.detach(), passes == 3)
    b_hi = torch.empty((B, K), device=dev, dtype=F32)
    b_lo = torch.empty((B, K), device=dev, dtype=F32) if passes == 3 else None
    call("regcn_gather_rows2", ptr(e_hi), ptr(e_lo), ptr(target), B, K, ptr(b_hi), ptr(b_lo))
    c, x2, y2, sm = (hyp[:4] if hyp is not None else (1.0, None, None, None))
    row_c = hyp[4] if hyp is not None and len(hyp) > 4 else None
    y2p = bp = None
    if hyp is not None or col_bias is not None:
        y2p = torch.empty(B, device=dev, dtype=F32) if hyp is not None else None
        bp = torch.empty(B, device=dev, dtype=F32) if col_bias is not None else None
        call("regcn_gather_scalars", None, ptr(y2), ptr(col_bias), ptr(target), ptr(target), B, None, ptr(y2p), ptr(bp))
    ts = torch.empty(B, device=dev, dtype=F32)
    call("regcn_pair_scores_tf32", ptr(q_hi), ptr(q_lo), ptr(b_hi), ptr(b_lo), B, K, int(hyp is not None), ptr(x2),
         ptr(y2p), ptr(bp), float(c), ptr(sm), ptr(row_c), ptr(ts), passes)
    nparts = _lib.load().regcn_score_lse_num_parts(N)
    pm = torch.empty(nparts * B, device=dev, dtype=F32)
    psum = torch.empty(nparts * B, device=dev, dtype=F32)
    call("regcn_score_lse_tf32", ptr(q_hi), ptr(q_lo), ptr(e_hi), ptr(e_lo), B, N, K, int(hyp is not None), ptr(x2),
         ptr(y2), ptr(col_bias), float(c), ptr(sm), ptr(row_c), passes, ptr(pm), ptr(psum))
    ce = torch.empty(B, device=dev, dtype=F32)
    loss = torch.empty(1, device=dev, dtype=F32)
    call("regcn_ce_from_lse", ptr(pm), ptr(psum), nparts, B, ptr(ts), ptr(ce), ptr(loss))
    return ce, loss


def ce_dense(score, triples, target_col):
    """(ce (B,), loss (1,)) of a materialised (B,N) score matrix (small candidate sets: relation prediction)."""
    B, N = score.shape
    ce = torch.empty(B, device=score.device, dtype=F32)
    loss = torch.empty(1, device=score.device, dtype=F32)
    call("regcn_ce_rows", score.data_ptr(), score.stride(0), B, N, ptr(triples), target_col, ptr(ce), ptr(loss))
    return ce, loss
