"""GPU: every kernel entry point of the C ABI against the CPU oracle (oracle/restate.py) on seeded inputs,
bit-exact for integer / index work, |a-b| <= 1e-4*max(1,|b|) for fp32, plus size-independent
properties at sizes the oracle cannot reach."""
import numpy as np
import pytest
import torch

from oracle import restate, synth
from tests.helpers import close

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    import regcn_b200
    from regcn_b200 import ops
    regcn_b200._lib.require_device()
    return regcn_b200, ops


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(DEV)


# ----------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("shape,zipf", [("tiny", True), ("small", True), ("c1", True), ("c3", False), ("c4d", True)])
def test_csr_build_bit_exact(shape, zipf):
    R, _ = _ops()
    case = synth.make_case(shape, 3, zipf=zipf)
    n, r = case["num_ents"], case["num_rels"]
    tri = case["history"][0]
    g = R.build_sub_graph(n, r, tri, True, 0)
    o = restate.build_edges(tri, n, r)
    E = 2 * len(tri)
    assert g.num_edges == E
    assert np.array_equal(g.src[:E].cpu().numpy(), o["src"])
    assert np.array_equal(g.dst[:E].cpu().numpy(), o["dst"])
    assert np.array_equal(g.etype[:E].cpu().numpy(), o["etype"])
    assert np.array_equal(g.indeg.cpu().numpy(), o["indeg"])
    assert np.array_equal(g.norm.cpu().numpy(), o["norm"])
    rowptr, eperm, src_sorted, etype_sorted = restate.csr_by_dst(o)
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.eperm[:E].cpu().numpy(), eperm)
    assert np.array_equal(g.src_sorted[:E].cpu().numpy(), src_sorted)
    assert np.array_equal(g.etype_sorted[:E].cpu().numpy(), etype_sorted)
    rel_rowptr, rel_ents = restate.r2e(tri, r)
    assert np.array_equal(g.rel_rowptr.cpu().numpy(), rel_rowptr)
    assert g.n_rel_ents == len(rel_ents)
    assert np.array_equal(g.rel_ents[:g.n_rel_ents].cpu().numpy(), rel_ents)
    # virtual rows: chunks of 32 edges per destination, in row order
    nch = -(-o["indeg"] // 32)                       # isolated destinations get no virtual row
    assert g.n_vrows == int(nch.sum())
    assert g.n_split_chunks == int(nch[nch > 1].sum())
    assert np.array_equal(g.vrow_row[:g.n_vrows].cpu().numpy(), np.repeat(np.arange(n), nch))
    active = np.nonzero(o["indeg"] > 0)[0]
    assert g.n_active == len(active)
    ap = np.full(n, -1, dtype=np.int64)
    ap[active] = np.arange(len(active))
    assert np.array_equal(g.active_pos.cpu().numpy(), ap)
    assert np.array_equal(g.active_rows[:g.n_active].cpu().numpy(), active)
    # the DGL-style views the reference's modules read
    assert g.number_of_nodes() == n and tuple(g.ndata["norm"].shape) == (n, 1)
    assert np.array_equal(g.edata["type"].cpu().numpy(), o["etype"])
    present = np.nonzero(np.diff(rel_rowptr))[0]
    assert np.array_equal(g.uniq_r, np.concatenate((present, present + r)))
    assert len(g.r_len) == 2 * len(present) and g.r_len[-1][1] == len(g.r_to_e)


def test_csr_build_empty_and_duplicates():
    R, ops = _ops()
    g = R.build_sub_graph(10, 3, np.zeros((0, 3), dtype=np.int64), True, 0)
    assert g.num_edges == 0 and g.n_vrows == 0 and g.n_split_chunks == 0 and g.n_rel_ents == 0 and g.n_active == 0
    assert np.array_equal(g.rowptr.cpu().numpy(), np.zeros(11))
    assert np.array_equal(g.norm.cpu().numpy(), np.ones(10, dtype=np.float32))
    h = torch.randn(10, 8, device=DEV)
    rel = torch.randn(6, 8, device=DEV)
    assert torch.count_nonzero(ops.union_aggregate(h, rel, g)) == 0      # DGL zero-fill for in-degree 0
    assert torch.count_nonzero(ops.rel_mean_pool(h, g)) == 0
    tri = np.array([[1, 0, 2], [1, 0, 2], [2, 1, 2]], dtype=np.int64)    # multi-edge + self loop are kept
    g = R.build_sub_graph(4, 2, tri, True, 0)
    assert g.indeg.cpu().tolist() == [0, 2, 4, 0]


def _graph_arrays(g):
    E = g.num_edges
    out = {k: getattr(g, k)[:E].cpu().numpy() for k in ("src", "dst", "etype", "src_sorted", "etype_sorted", "eperm")}
    out.update({k: getattr(g, k).cpu().numpy() for k in ("indeg", "norm", "rowptr", "vptr", "sptr", "active_pos",
                                                         "rel_rowptr")})
    out["vrow_row"] = g.vrow_row[:g.n_vrows].cpu().numpy()
    out["rel_ents"] = g.rel_ents[:g.n_rel_ents].cpu().numpy()
    out["active_rows"] = g.active_rows[:g.n_active].cpu().numpy()
    out["counts"] = np.array([g.n_vrows, g.n_split_chunks, g.n_rel_ents, g.max_hub_degree, g.n_active])
    return out


def test_csr_build_batch_matches_single_builds():
    """regcn_csr_build_batch (one CTA per small snapshot, every sort width, plus the CUB path for a big one and an
    empty snapshot in the same call) is bit-identical to regcn_csr_build and to the oracle."""
    R, _ = _ops()
    from regcn_b200.graph import build_sub_graphs
    rng = np.random.default_rng(5)
    n, r = 3000, 37
    sizes = [0, 1, 250, 2048, 2049, 3500, 4096, 6000, 8192, 9000]      # E = 2T: widths 4 / 8 / 16 and > 16384 edges
    snaps = []
    for i, t in enumerate(sizes):
        zipf = i % 2 == 0
        snaps.append(synth.make_snapshot(rng, n, r, t, zipf) if t else np.zeros((0, 3), dtype=np.int64))
    gs = build_sub_graphs(n, r, [torch.from_numpy(s) for s in snaps], torch.device(DEV))
    for tri, g in zip(snaps, gs):
        single = R.build_sub_graph(n, r, tri, True, 0)
        a, b = _graph_arrays(g), _graph_arrays(single)
        for k in a:
            assert np.array_equal(a[k], b[k]), (len(tri), k)
        o = restate.build_edges(tri, n, r)
        rowptr, eperm, src_sorted, etype_sorted = restate.csr_by_dst(o)
        assert np.array_equal(a["rowptr"], rowptr) and np.array_equal(a["eperm"], eperm)
        assert np.array_equal(a["src_sorted"], src_sorted) and np.array_equal(a["etype_sorted"], etype_sorted)
        rel_rowptr, rel_ents = restate.r2e(tri, r)
        assert np.array_equal(a["rel_rowptr"], rel_rowptr) and np.array_equal(a["rel_ents"], rel_ents)


def test_csr_build_batch_many_snapshots_and_big_ids():
    """More snapshots than one launch holds (16), and entity / relation counts near the packed-key limit."""
    R, _ = _ops()
    from regcn_b200.graph import build_sub_graphs
    rng = np.random.default_rng(6)
    n, r = 23033, 256
    snaps = [synth.make_snapshot(rng, n, r, 100 + 37 * i, True) for i in range(19)]
    gs = build_sub_graphs(n, r, [torch.from_numpy(s) for s in snaps], torch.device(DEV))
    for tri, g in zip(snaps, gs):
        a, b = _graph_arrays(g), _graph_arrays(R.build_sub_graph(n, r, tri, True, 0))
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    n, r = 262_144, 8191                                               # largest N / R the one-CTA path takes
    tri = np.stack((rng.integers(0, n, 500), rng.integers(0, r, 500), rng.integers(0, n, 500)), 1).astype(np.int64)
    tri[:8] = [[n - 1, r - 1, n - 1]] * 4 + [[0, 0, n - 1]] * 4
    g = build_sub_graphs(n, r, [torch.from_numpy(tri)], torch.device(DEV))[0]
    a, b = _graph_arrays(g), _graph_arrays(R.build_sub_graph(n, r, tri, True, 0))
    for k in a:
        assert np.array_equal(a[k], b[k]), k


# ----------------------------------------------------------------------------------------- K4 / K2
@pytest.mark.parametrize("impl", [1, 2, 3])
@pytest.mark.parametrize("shape,d,radius", [("tiny", 200, False), ("small", 200, True), ("c1", 200, False),
                                            ("c4d", 200, True), ("small", 64, False), ("c4d", 128, False)])
def test_union_aggregate_vs_oracle(shape, d, radius, impl):
    """impl 1 = register-staged gathers, impl 2 = cp.async.bulk gathers staged in shared memory."""
    R, ops = _ops()
    R._lib.load().regcn_aggregate_tune(impl)
    case = synth.make_case(shape, 5)
    n, r = case["num_ents"], case["num_rels"]
    tri = case["history"][0]
    g = R.build_sub_graph(n, r, tri, True, 0)
    o = restate.build_edges(tri, n, r)
    rng = np.random.default_rng(1)
    h = rng.standard_normal((n, d)).astype(np.float32)
    rel = rng.standard_normal((2 * r, d)).astype(np.float32)
    rad = rng.uniform(0.5, 3.0, n).astype(np.float32) if radius else None
    gamma = 0.15
    out = ops.union_aggregate(_t(h), _t(rel), g, radius=_t(rad) if radius else None, gamma=gamma).cpu().numpy()
    msg = h[o["src"]].astype(np.float64) + rel[o["etype"]]
    if radius:
        w32 = np.exp(-np.float32(gamma) * np.abs(rad[o["src"]] - rad[o["dst"]])).astype(np.float64)
        msg = msg * w32[:, None]
    ref = np.zeros((n, d))
    np.add.at(ref, o["dst"], msg)
    ref *= o["norm"][:, None]
    # hub rows sum thousands of O(1) terms in fp32: gate relative to the row's accumulated magnitude
    mag = np.zeros(n)
    np.add.at(mag, o["dst"], np.abs(msg).max(axis=1))
    scale = np.maximum(1.0, (mag * o["norm"])[:, None])
    R._lib.load().regcn_aggregate_tune(0)
    assert np.all(np.abs(out - ref) <= 1e-4 * np.maximum(scale, np.abs(ref))), np.abs(out - ref).max()
    if g.n_split_chunks:
        assert g.max_hub_degree > 32


def test_union_aggregate_large_vs_torch():
    """Full-size property (oracle-free): 200k entities, 2M edges against torch.index_add_ in fp64 on the GPU."""
    R, ops = _ops()
    n, r, t, d = 200_000, 64, 1_000_000, 200
    rng = np.random.default_rng(0)
    tri = synth.make_snapshot(rng, n, r, t, zipf=True)
    g = R.build_sub_graph(n, r, tri, True, 0)
    h = torch.randn(n, d, device=DEV)
    rel = torch.randn(2 * r, d, device=DEV)
    out = ops.union_aggregate(h, rel, g)
    E = g.num_edges
    msg = h[g.src[:E].long()].double() + rel[g.etype[:E].long()].double()
    ref = torch.zeros(n, d, device=DEV, dtype=torch.float64).index_add_(0, g.dst[:E].long(), msg)
    ref *= g.norm.double().view(-1, 1)
    err = (out.double() - ref).abs()
    assert bool((err <= 1e-4 * torch.clamp(ref.abs(), min=1.0)).all()), float(err.max())
    # linearity in h (size-independent property): agg(h1 + h2, rel) = agg(h1, rel) + agg(h2, 0)
    h2 = torch.randn(n, d, device=DEV)
    lhs = ops.union_aggregate(h + h2, rel, g)
    rhs = out + ops.union_aggregate(h2, torch.zeros_like(rel), g)
    assert bool(((lhs - rhs).abs() <= 1e-3 * torch.clamp(lhs.abs(), min=1.0)).all())


@pytest.mark.parametrize("shape,nsplit", [("tiny", 1), ("c1", 1), ("c4d", 1), ("c4d", 7)])
def test_rel_mean_pool_vs_oracle(shape, nsplit):
    R, ops = _ops()
    case = synth.make_case(shape, 2)
    n, r = case["num_ents"], case["num_rels"]
    tri = case["history"][1]
    g = R.build_sub_graph(n, r, tri, True, 0)
    h = torch.randn(n, 200, dtype=torch.float64)
    rel_rowptr, rel_ents = restate.r2e(tri, r)
    ref = restate.rel_mean_pool(h, rel_rowptr, rel_ents, r).numpy()
    out = ops.rel_mean_pool(h.float().to(DEV), g, nsplit=nsplit).cpu().numpy()
    ok, worst = close(out, ref)
    assert ok, worst


# ----------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,trans_b,bias,split_k", [
    (1, 200, 200, False, False, 1), (513, 200, 200, False, True, 1), (300, 600, 400, True, True, 1),
    (1000, 23033 // 8, 200, True, False, 1), (257, 200, 10000, True, True, 6), (64, 8, 12, False, False, 1),
    (130, 131, 204, True, True, 3), (2914, 200, 200, False, False, 1)])
def test_gemm_f32_vs_fp64(M, N, K, trans_b, bias, split_k):
    _, ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g)
    b = torch.randn((N, K) if trans_b else (K, N), generator=g)
    bv = torch.randn(N, generator=g) if bias else None
    ref = a.double() @ (b.double().t() if trans_b else b.double())
    if bias:
        ref = ref + bv.double()
    out = ops.gemm(a.to(DEV), b.to(DEV), trans_b=trans_b, bias=bv.to(DEV) if bias else None, split_k=split_k)
    # fp32 accumulation over K terms of unit variance: error ~ sqrt(K)*eps*|terms|
    tol = 1e-4 * max(1.0, (K / 200.0) ** 0.5)
    err = (out.cpu().double() - ref).abs()
    assert bool((err <= tol * torch.clamp(ref.abs(), min=1.0) * 4).all()), float(err.max())
    # accumulate into a strided output block
    big = torch.zeros(M, N + 8, device=DEV)
    view = big[:, 4:4 + N]
    view.copy_(torch.ones(M, N))
    ops.gemm(a.to(DEV), b.to(DEV), trans_b=trans_b, out=view, accumulate=True, split_k=split_k)
    ref2 = a.double() @ (b.double().t() if trans_b else b.double()) + 1.0
    err2 = (view.cpu().double() - ref2).abs()
    assert bool((err2 <= tol * torch.clamp(ref2.abs(), min=1.0) * 4).all())
    assert float(big[:, :4].abs().sum()) == 0.0 and float(big[:, 4 + N:].abs().sum()) == 0.0


def test_gemm_f32_yardstick_vs_fp64():
    """regcn_gemm_f32 (fp32 CUDA cores): kept as the yardstick the tensor-core path is compared with."""
    _, ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(11)
    a, b = torch.randn(130, 204, generator=g), torch.randn(131, 204, generator=g)
    out = ops.gemm_f32_yardstick(a.to(DEV), b.to(DEV), trans_b=True, split_k=3)
    ref = a.double() @ b.double().t()
    assert float((out.cpu().double() - ref).abs().max()) <= 4e-4
    tc = ops.gemm(a.to(DEV), b.to(DEV), trans_b=True)
    assert float((tc - out).abs().max()) <= 4e-4


def test_gemm_rejects_bad_arguments():
    _, ops = _ops()
    a = torch.randn(4, 6, device=DEV)          # K = 6 is not a multiple of 4
    b = torch.randn(6, 8, device=DEV)
    with pytest.raises(ValueError, match="multiple of 4"):
        ops.gemm(a, b)                          # one backend: no silent route to the CUDA-core kernel
    with pytest.raises(ValueError):
        ops.set_gemm_impl("simt")               # the fp32 CUDA-core kernel is a test yardstick, not a product backend
    with pytest.raises(RuntimeError):
        ops.gemm(torch.randn(4, 8), torch.randn(8, 8))   # CPU tensors are refused: no fallback


# ----------------------------------------------------------------------------------------- row maps
@pytest.mark.parametrize("d", [200, 64, 256])
def test_row_maps_vs_oracle(d):
    _, ops = _ops()
    c = 0.01
    g = torch.Generator().manual_seed(d)
    x = torch.randn(301, d, generator=g) * torch.logspace(-3, 1.2, 301).view(-1, 1)   # tiny to beyond the ball
    x[7] = 0
    xd = x.to(DEV)
    pairs = [
        (ops.ROW_NORMALIZE, restate.normalize_rows(x)),
        (ops.ROW_TANH, torch.tanh(x)),
        (ops.ROW_EXP0, restate.exp0(x, c)),
        (ops.ROW_PROJECT, restate.project(x, c)),
    ]
    ball = restate.project(x * 0.2, c)
    for mode, ref in pairs:
        ok, worst = close(ops.row_map(xd, mode, c=c).cpu().numpy(), ref.numpy())
        assert ok, (mode, worst)
    bd = ball.to(DEV)
    t = restate.log0(ball, c)
    for mode, ref in [(ops.ROW_LOG0, t), (ops.ROW_LEAKY_TANH_LOG0, 0.9 * torch.tanh(t) + 0.1 * t),
                      (ops.ROW_TANGENT_NORMALIZE, restate.exp0(restate.normalize_rows(t), c))]:
        out, ss = ops.row_map(bd, mode, c=c, want_sumsq=True)
        ok, worst = close(out.cpu().numpy(), ref.numpy())
        assert ok, (mode, worst)
        ok, worst = close(ss.cpu().numpy(), (ref.double() ** 2).sum(1).numpy())
        assert ok, (mode, worst)
    ok, worst = close(ops.row_sumsq(xd).cpu().numpy(), (x.double() ** 2).sum(1).numpy())
    assert ok, worst


def test_gru_gate_vs_oracle():
    _, ops = _ops()
    d, M = 200, 460
    g = torch.Generator().manual_seed(0)
    x, h = torch.randn(M, 2 * d, generator=g), torch.randn(M, d, generator=g)
    w_ih, w_hh = torch.randn(3 * d, 2 * d, generator=g) * 0.05, torch.randn(3 * d, d, generator=g) * 0.05
    b_ih, b_hh = torch.randn(3 * d, generator=g) * 0.1, torch.randn(3 * d, generator=g) * 0.1
    ref = restate.gru_cell(x.double(), h.double(), w_ih.double(), w_hh.double(), b_ih.double(), b_hh.double())
    torch_ref = torch.nn.functional.normalize(ref.float())
    gi = ops.gemm(x.to(DEV), w_ih.to(DEV), trans_b=True, bias=b_ih.to(DEV))
    gh = ops.gemm(h.to(DEV), w_hh.to(DEV), trans_b=True, bias=b_hh.to(DEV))
    ok, worst = close(ops.gru_gate(gi, gh, h.to(DEV), False).cpu().numpy(), ref.numpy())
    assert ok, worst
    ok, worst = close(ops.gru_gate(gi, gh, h.to(DEV), True).cpu().numpy(), torch_ref.numpy())
    assert ok, worst
    # cross-check the restated cell against torch's own GRUCell
    cell = torch.nn.GRUCell(2 * d, d)
    with torch.no_grad():
        cell.weight_ih.copy_(w_ih); cell.weight_hh.copy_(w_hh); cell.bias_ih.copy_(b_ih); cell.bias_hh.copy_(b_hh)
        ok, worst = close(cell(x, h).numpy(), ref.numpy())
    assert ok, worst


def test_union_combine_and_time_gate_vs_oracle():
    _, ops = _ops()
    n, d, c = 777, 200, 0.01
    g = torch.Generator().manual_seed(4)
    P = torch.randn(n, d, generator=g) * 6          # exercises the +-10 clamps
    L = torch.randn(n, 2 * d, generator=g) * 3
    indeg = (torch.rand(n, generator=g) > 0.4).int()
    S = torch.randn(n, d, generator=g)
    sb = torch.randn(d, generator=g) * 0.1
    prev = torch.randn(n, d, generator=g)
    loop = torch.where(indeg.view(-1, 1) > 0, L[:, :d], L[:, d:])
    dev = lambda t: t.to(DEV)
    # Euclidean
    out, _, _ = ops.union_combine(dev(P), dev(L), dev(indeg), act=1)
    ok, worst = close(out.cpu().numpy(), restate.rrelu(P + loop).numpy())
    assert ok, worst
    sw = torch.sigmoid(S + sb)
    out, _, _ = ops.union_combine(dev(P), dev(L), dev(indeg), act=1, skip=dev(S), skip_bias=dev(sb), prev=dev(prev))
    ok, worst = close(out.cpu().numpy(), restate.rrelu(sw * (P + loop) + (1 - sw) * prev).numpy())
    assert ok, worst
    out, _, _ = ops.union_combine(dev(P), None, None, act=0)
    assert torch.equal(out.cpu(), P)
    # hyperbolic: clamp, loop, clamp, rrelu, exp_0 (+ tangent / radius of the result)
    ref_t = restate.rrelu((P.clamp(-10, 10) + loop).clamp(-10, 10))
    ref_h = restate.exp0(ref_t, c)
    out, ht, rad = ops.union_combine(dev(P), dev(L), dev(indeg), act=1, hyper=True, c=c, want_tangent=True,
                                     want_radius=True)
    ok, worst = close(out.cpu().numpy(), ref_h.numpy())
    assert ok, worst
    # tangent / radius hand-over to the next layer, on points well inside the ball (at the boundary atanh has a
    # condition number of ~5e5 and the reference's own fp32 result is noise-limited)
    P2, L2 = P * 0.03, L * 0.03
    loop2 = torch.where(indeg.view(-1, 1) > 0, L2[:, :d], L2[:, d:])
    ref_h2 = restate.exp0(restate.rrelu(P2 + loop2), c)
    out, ht, rad = ops.union_combine(dev(P2), dev(L2), dev(indeg), act=1, hyper=True, c=c, want_tangent=True,
                                     want_radius=True)
    ok, worst = close(out.cpu().numpy(), ref_h2.numpy())
    assert ok, worst
    ok, worst = close(ht.cpu().numpy(), restate.log0(ref_h2, c).numpy())
    assert ok, worst
    ok, worst = close(rad.cpu().numpy(), restate.get_radius(ref_h2).numpy())
    assert ok, worst
    # Euclidean time gate
    G, b, cur, h = torch.randn(n, d, generator=g), torch.randn(d, generator=g), P, prev
    for ln in (False, True):
        tw = torch.sigmoid(G + b)
        ref = tw * (restate.normalize_rows(cur) if ln else cur) + (1 - tw) * h
        ok, worst = close(ops.time_gate(dev(G), dev(b), dev(cur), dev(h), ln).cpu().numpy(), ref.numpy())
        assert ok, worst


@pytest.mark.parametrize("layer_norm,residual", [(False, True), (True, True), (False, False)])
def test_hyperbolic_row_kernels_vs_oracle(layer_norm, residual):
    _, ops = _ops()
    n, d, c = 513, 200, 0.01
    rmin, rmax, beta, eps_r = 0.5, 3.0, 0.7, 0.1
    g = torch.Generator().manual_seed(9)
    emb = torch.randn(n, d, generator=g)
    rs_raw = torch.rand(n, generator=g) * 4.0          # some outside [rmin, rmax]
    rs = restate.static_radius(rs_raw, c, rmin, rmax)
    init = restate.normalize_rows(emb) if layer_norm else emb
    ref_h = restate.apply_radius(restate.exp0(init, c), rs, c)
    h = ops.hyp_init(emb.to(DEV), rs_raw.to(DEV), layer_norm, False, c, rmin, rmax)
    ok, worst = close(h.cpu().numpy(), ref_h.numpy())
    assert ok, worst
    ht, pt, rad = ops.hyp_tangent(h, c)
    ok, worst = close(ht.cpu().numpy(), restate.log0(ref_h, c).numpy())
    assert ok, worst
    ok, worst = close(rad.cpu().numpy(), restate.get_radius(ref_h).numpy())
    assert ok, worst
    # time gate + radius evolution
    h2 = restate.exp0(torch.randn(n, d, generator=g) * 3, c)
    G = torch.randn(n, d, generator=g)
    b = torch.randn(d, generator=g) * 0.1
    w = torch.randn(1, d, generator=g) * 0.05
    rb = 0.03
    cur = restate.project(h2, c)
    if layer_norm:
        cur = restate.exp0(restate.normalize_rows(restate.log0(cur, c)), c)
    ct = restate.log0(cur, c).clamp(-10, 10)
    ptr_ = restate.log0(ref_h, c).clamp(-10, 10)
    tw = torch.sigmoid(G + b)
    hn = restate.project(restate.exp0(tw * ct + (1 - tw) * ptr_, c), c)
    if residual:
        hn = restate.radius_evolution(hn, rs, w, torch.tensor([rb]), c, beta, eps_r)
    else:
        hn = restate.apply_radius(hn, rs, c)
    out = ops.hyp_time_gate(h2.to(DEV), pt, G.to(DEV), b.to(DEV), rs_raw.to(DEV), w.view(-1).to(DEV), rb, layer_norm,
                            residual, c, rmin, rmax, beta, eps_r)
    ok, worst = close(out.cpu().numpy(), hn.numpy(), rtol=1e-4)
    assert ok, worst


# ----------------------------------------------------------------------------------------- K6 / K7
@pytest.mark.parametrize("nb", [100, 10, 50])
def test_block_and_lorentz_aggregate_vs_oracle(nb):
    R, ops = _ops()
    case = synth.make_case("small_l", 4)
    n, r, d, c = case["num_ents"], case["num_rels"], 200, 0.01
    tri = case["history"][2]
    g = R.build_sub_graph(n, r, tri, True, 0)
    o = restate.build_edges(tri, n, r)
    gen = torch.Generator().manual_seed(nb)
    sb = d // nb
    W = torch.randn(2 * r, nb * sb * sb, generator=gen) * 0.3
    rel = torch.randn(2 * r, d, generator=gen) * 0.3
    h = restate.exp0(torch.randn(n, d, generator=gen), c)
    # K6 (static-graph layer): rrelu(norm * sum blockdiag(W).h) with the activation done by union_combine
    ref = restate.block_layer(h.double(), o, W.double(), nb, d)
    agg = ops.block_aggregate(h.to(DEV), W.to(DEV), g, nb, d)
    out, _, _ = ops.union_combine(agg, None, None, act=1)
    ok, worst = close(out.cpu().numpy(), ref.numpy())
    assert ok, worst
    # K7: Lorentz centroid aggregate -> full layer output
    wl, we = torch.randn(d, d, generator=gen) * 0.1, torch.randn(d, d, generator=gen) * 0.1
    ref_l = restate.lorentz_layer(h, rel, o, W, wl, we, c, nb)
    ht, _, _ = ops.hyp_tangent(h.to(DEV), c, want_clamped=False, want_radius=False)
    agg = ops.lorentz_aggregate(ht, W.to(DEV), rel.to(DEV), g, nb, c)
    L = ops.gemm(ht, torch.cat((wl, we), 1).to(DEV))
    out, _, _ = ops.union_combine(agg, L, g.indeg, act=1, hyper=True, c=c)
    ok, worst = close(out.cpu().numpy(), ref_l.numpy(), rtol=1e-4)
    assert ok, worst


# ----------------------------------------------------------------------------------------- decoders
def test_convtranse_tower_and_hyp_query_vs_oracle():
    R, ops = _ops()
    from tests.helpers import build_model
    cfg = dict(kind="regcn", layer_norm=True, seed=11)
    n, r, B, d = 300, 20, 77, 200
    m, sd = build_model(cfg, n, r)
    m = m.to(DEV)
    gen = torch.Generator().manual_seed(2)
    emb = torch.randn(n, d, generator=gen)
    rel = torch.randn(2 * r, d, generator=gen)
    tri = torch.stack([torch.randint(0, n, (B,), generator=gen), torch.randint(0, 2 * r, (B,), generator=gen),
                       torch.randint(0, n, (B,), generator=gen)], 1)
    P = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    ref = restate.convtranse_scores(P, emb, rel, tri.numpy())
    out = m.decoder_ob(emb.to(DEV), rel.to(DEV), tri.to(DEV), mode="test")
    ok, worst = close(out.cpu().numpy(), ref.numpy(), rtol=1e-4)
    assert ok, worst
    ref = restate.convtransr_scores(P, emb, rel, tri.numpy())
    out = m.rdecoder(emb.to(DEV), rel.to(DEV), tri.to(DEV), mode="test")
    ok, worst = close(out.cpu().numpy(), ref.numpy(), rtol=1e-4)
    assert ok, worst
    # single-query batch skips bn2 (src/decoder.py:93-94)
    ref1 = restate.convtranse_scores(P, emb, rel, tri[:1].numpy())
    out1 = m.decoder_ob(emb.to(DEV), rel.to(DEV), tri[:1].contiguous().to(DEV), mode="test")
    ok, worst = close(out1.cpu().numpy(), ref1.numpy(), rtol=1e-4)
    assert ok, worst


@pytest.mark.parametrize("B,n,d,C,nout,cols", [(77, 300, 200, 50, 200, (0, 1)), (2914, 5000, 200, 50, 200, (0, 2)),
                                               (1, 40, 200, 50, 200, (0, 1)), (300, 64, 72, 7, 40, (2, 1)),
                                               (129, 64, 16, 64, 16, (0, 2))])
def test_convtrans_fc_fused_tower_vs_fp64_and_feature_map_path(B, n, d, C, nout, cols):
    """regcn_convtrans_fc (bn0 -> conv1d -> bn1 -> relu -> fc with the feature map computed inside the GEMM's operand ring)
    against an fp64 evaluation of src/decoder.py:81-91 and against the path that writes the feature map
    (regcn_convtranse_features + the fp32-A GEMM): same element arithmetic, another summation order of the FC.  Row tails
    (B % 128), a last position block of 8 valid columns (d = 200, 72), d = 16 (one block) and 64 channels are covered."""
    R, ops = _ops()
    import torch.nn.functional as F_
    gen = torch.Generator().manual_seed(B + d)
    ent = torch.randn(n, d, generator=gen)
    sec = torch.randn(n, d, generator=gen)
    tri = torch.stack([torch.randint(0, n, (B,), generator=gen) for _ in range(3)], 1)
    bn0 = (torch.rand(2, generator=gen) + 0.5, torch.randn(2, generator=gen) * 0.1)
    bn1 = (torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.1)
    cw = torch.randn(C, 2, 3, generator=gen) * 0.4
    cb = torch.randn(C, generator=gen) * 0.1
    fw = torch.nn.Parameter(torch.randn(nout, C * d, generator=gen) / (C * d) ** 0.5)
    fb = torch.randn(nout, generator=gen) * 0.1
    x = torch.stack([ent[tri[:, cols[0]]], sec[tri[:, cols[1]]]], 1).double()
    x = x * bn0[0].double()[None, :, None] + bn0[1].double()[None, :, None]
    y = F_.conv1d(x, cw.double(), cb.double(), padding=1)
    y = torch.relu(y * bn1[0].double()[None, :, None] + bn1[1].double()[None, :, None]).reshape(B, -1)
    ref = y @ fw.detach().double().t() + fb.double()
    cu = lambda t: t.to(DEV).contiguous()
    fw_d = torch.nn.Parameter(cu(fw.detach()))
    bn0d, bn1d = (cu(bn0[0]), cu(bn0[1])), (cu(bn1[0]), cu(bn1[1]))
    assert ops.convtrans_fc_ok(d, cw, nout)
    out = ops.convtrans_fc(cu(ent), cu(sec), cu(tri), cols[0], cols[1], bn0d, cu(cw), cu(cb), bn1d, fw_d, cu(fb))
    out2 = ops.convtrans_fc(cu(ent), cu(sec), cu(tri), cols[0], cols[1], bn0d, cu(cw), cu(cb), bn1d, fw_d, cu(fb))
    assert torch.equal(out, out2)                                     # deterministic (fixed split order)
    feats = ops.convtranse_features(cu(ent), cu(sec), cu(tri), cols[0], cols[1], bn0d, cu(cw), cu(cb), bn1d, split=False)
    old = ops.gemm(feats, fw_d.detach(), trans_b=True, bias=cu(fb), split_k=4 if C * d >= 2048 else 1, split_a_on_chip=True)
    scale = float(ref.abs().max())
    err_new = float((out.double().cpu() - ref).abs().max()) / scale
    err_old = float((old.double().cpu() - ref).abs().max()) / scale
    assert err_new <= max(2e-6, 2.0 * err_old), (err_new, err_old)
    # the tower's tail (bn2 -> relu) folded into the split-K reduction
    bn2 = (cu(torch.rand(nout, generator=gen) + 0.5), cu(torch.randn(nout, generator=gen) * 0.1))
    tail = ops.convtrans_fc(cu(ent), cu(sec), cu(tri), cols[0], cols[1], bn0d, cu(cw), cu(cb), bn1d, fw_d, cu(fb), bn2=bn2,
                            relu=True)
    want = torch.relu(out.double() * bn2[0].double() + bn2[1].double())
    assert float((tail.double() - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))
    if B > 1:
        # a slice of a sharded batch splits like the whole batch: rows do not depend on the cut
        part = ops.convtrans_fc(cu(ent), cu(sec), cu(tri[:B // 2]), cols[0], cols[1], bn0d, cu(cw), cu(cb), bn1d, fw_d, cu(fb),
                                batch_total=B)
        assert torch.equal(part, out[:B // 2])


@pytest.mark.parametrize("decoder", ["roth", "murp", "hyperbolic_convtranse"])
def test_hyperbolic_decoders_vs_oracle(decoder):
    R, ops = _ops()
    from tests.helpers import build_model
    c = 0.01
    cfg = dict(kind="hyp", layer_norm=False, seed=5, decoder=decoder, encoder="hyperbolic_uvrgcn", gamma=0.15)
    n, r, B, d = 400, 25, 90, 200
    m, sd = build_model(cfg, n, r)
    m = m.to(DEV)
    gen = torch.Generator().manual_seed(3)
    emb = restate.apply_radius(restate.exp0(torch.randn(n, d, generator=gen), c), torch.rand(n, generator=gen) * 2.5 + 0.5, c)
    rel = torch.randn(2 * r, d, generator=gen) * 0.3
    tri = torch.stack([torch.randint(0, n, (B,), generator=gen), torch.randint(0, 2 * r, (B,), generator=gen),
                       torch.randint(0, n, (B,), generator=gen)], 1)
    P = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    fn = {"roth": (restate.roth_scores, restate.rothrel_scores), "murp": (restate.murp_scores, restate.murprel_scores),
          "hyperbolic_convtranse": (restate.hyp_convtranse_scores, restate.hyp_convtransr_scores)}[decoder]
    ref = fn[0](P, emb, rel, tri.numpy(), c)
    ref = ref[0] if isinstance(ref, tuple) else ref
    out = m.decoder_ob(emb.to(DEV), rel.to(DEV), tri.to(DEV), mode="test")
    ok, worst = close(out.cpu().numpy(), ref.numpy(), rtol=1e-4)
    assert ok, worst
    ref = fn[1](P, emb, rel, tri.numpy(), c)
    ref = ref[0] if isinstance(ref, tuple) else ref
    out = m.rdecoder(emb.to(DEV), rel.to(DEV), tri.to(DEV), mode="test")
    ok, worst = close(out.cpu().numpy(), ref.numpy(), rtol=1e-4)
    assert ok, worst


# ----------------------------------------------------------------------------------------- K14
@pytest.mark.parametrize("B,N,levels", [(50, 300, 0), (200, 7128, 0), (64, 1000, 7), (1, 5, 2)])
def test_rank_bit_exact_vs_oracle(B, N, levels):
    R, ops = _ops()
    from regcn_b200 import utils
    rng = np.random.default_rng(B + N)
    score = rng.standard_normal((B, N)).astype(np.float32)
    if levels:
        score = np.round(score * levels) / levels        # heavy ties, including with the target
    r = 5
    tri = np.stack([rng.integers(0, 40, B), rng.integers(0, r, B), rng.integers(0, min(N, 40), B)], 1).astype(np.int64)
    all_ans = {}
    for h, rr, t in tri:
        all_ans.setdefault(int(h), {}).setdefault(int(rr), set()).add(int(t))
    for h in all_ans:                                     # add extra true answers so the filter has work to do
        for rr in all_ans[h]:
            all_ans[h][rr].update(int(x) for x in rng.integers(0, N, 4))
    fm, m, rank, frank = restate.total_rank(tri, score, all_ans, 0)
    s_dev = _t(score)
    fm2, m2, rank2, frank2 = utils.get_total_rank(_t(tri, torch.int64), s_dev, all_ans, 1000, rel_predict=0)
    assert np.array_equal(rank2.cpu().numpy(), rank)
    assert np.array_equal(frank2.cpu().numpy(), frank)
    assert abs(fm - fm2) < 1e-6 and abs(m - m2) < 1e-6
    # the reference's in-place side effect on the score matrix (utils.py:60)
    assert np.array_equal(s_dev.cpu().numpy(), restate.filter_scores(tri, score, all_ans, 0))
    # relation-prediction flavour (target = column 1, answers keyed by (h, t))
    score_r = rng.standard_normal((B, 2 * r)).astype(np.float32)
    all_ans_r = {}
    for h, rr, t in tri:
        all_ans_r.setdefault(int(h), {}).setdefault(int(t), set()).add(int(rr))
    fm, m, rank, frank = restate.total_rank(tri, score_r, all_ans_r, 1)
    _, _, rank2, frank2 = utils.get_total_rank(_t(tri, torch.int64), _t(score_r), all_ans_r, 1000, rel_predict=1)
    assert np.array_equal(rank2.cpu().numpy(), rank) and np.array_equal(frank2.cpu().numpy(), frank)
    # all_ans=None: filtered == raw
    _, _, rank3, frank3 = utils.get_total_rank(_t(tri, torch.int64), _t(score), None, 1000, rel_predict=0)
    assert torch.equal(rank3, frank3)


def test_filter_csr_from_snapshot_matches_dict():
    R, _ = _ops()
    from regcn_b200 import utils
    case = synth.make_case("small", 8)
    r = case["num_rels"]
    all_t = restate.add_inverse(case["test"], r)
    for rel_p, nk in ((0, 2 * r), (1, case["num_ents"])):
        d = synth.answers_of(case["test"], r, bool(rel_p))
        a = utils.filter_csr_from_dict(torch.as_tensor(all_t), d, rel_predict=rel_p, device=DEV)
        b = utils.filter_csr_from_snapshot(_t(all_t, torch.int64), nk, rel_predict=rel_p)
        assert a.lists() == b.lists()


def test_filter_lists_hub_queries_and_duplicates():
    """Warp-per-query filter kernels: lists longer than a warp (hub subject with > 32 answers), duplicate triples,
    a single query, and the pair lists of the fused rank path."""
    R, _ = _ops()
    from regcn_b200 import utils
    rng = np.random.default_rng(11)
    n, r = 500, 7
    hub = np.stack((np.full(90, 3), np.full(90, 2), rng.integers(0, n, 90)), 1)        # 90 answers, some repeated
    dup = np.array([[5, 1, 9]] * 4 + [[5, 1, 8]] * 2)
    rest = np.stack((rng.integers(0, n, 300), rng.integers(0, r, 300), rng.integers(0, n, 300)), 1)
    # list lengths around the register (32), warp-rank-sort (256) and serial paths, in a query set larger than one
    # shared-memory key tile (2048): hubs with 32 / 33 / 256 / 300 distinct answers plus repeats
    sized = [np.stack((np.full(k + 5, 10 + i), np.full(k + 5, i % r), np.concatenate((np.arange(k) * 3 % n if k <= n // 3 else np.arange(k), rng.integers(0, 4, 5)))), 1)
             for i, k in enumerate((32, 33, 256, 300))]
    big = np.concatenate(sized + [np.stack((rng.integers(0, n, 1500), rng.integers(0, r, 1500), rng.integers(0, n, 1500)), 1)])
    for test in (np.concatenate((hub, dup, rest)).astype(np.int64), np.array([[1, 0, 2]], dtype=np.int64), big.astype(np.int64)):
        all_t = restate.add_inverse(test, r)
        for rel_p in (0, 1):
            d = synth.answers_of(test, r, bool(rel_p))
            a = utils.filter_csr_from_dict(torch.as_tensor(all_t), d, rel_predict=rel_p, device=DEV)
            b = utils.filter_lists_from_queries(_t(all_t, torch.int64), rel_p)
            assert a.lists() == b.lists()
            pfs = [utils.filter_lists_begin(_t(all_t, torch.int64), k) for k in (0, 1)]
            pfs[1].triples = pfs[0].triples
            two = utils.filter_lists_finish2(pfs[0], int(pfs[0].total), pfs[1], int(pfs[1].total))[rel_p]
            assert two.lists() == b.lists() and all(torch.equal(x, y) for x, y in zip(two.pairs(None), b.pairs(None)))
            B = all_t.shape[0]
            pa, pe = b.pairs(None)
            assert pa[:B].cpu().tolist() == list(range(B))
            assert pe[:B].cpu().tolist() == all_t[:, 1 if rel_p else 2].tolist()
            beg = b.ptr.cpu().numpy()
            rows = pa[B:].cpu().numpy()
            cand = pe[B:].cpu().numpy()
            lists = b.lists()
            for q in rng.integers(0, B, 20):
                sl = slice(beg[q], beg[q] + int((rows == q).sum()))
                assert (rows[sl] == q).all() and set(cand[sl].tolist()) == set(lists[q])


def test_queries_prepare_equals_stepwise_preparation():
    """regcn_queries_prepare (inverse triples + both filter count passes + their scans in one call) against the framework
    operations it replaces (flip / add / cat, regcn_filter_count + cumsum), incl. more queries than one 1024-thread scan
    slice per thread and a single triple."""
    R, _ = _ops()
    from regcn_b200 import utils
    rng = np.random.default_rng(3)
    for T, n, r in ((1, 50, 4), (37, 50, 4), (1457, 23033, 256), (5000, 300, 7)):
        tri = np.stack((rng.integers(0, n, T), rng.integers(0, r, T), rng.integers(0, n, T)), 1).astype(np.int64)
        t = torch.from_numpy(tri).to(DEV)
        all_t, pf_e, pf_r, totals = utils.queries_prepare(t, r)
        ref_all = torch.from_numpy(restate.add_inverse(tri, r)).to(DEV)
        assert torch.equal(all_t, ref_all)
        for pf, rel_p in ((pf_e, 0), (pf_r, 1)):
            old = utils.filter_lists_begin(ref_all, rel_p)
            assert torch.equal(pf.beg, old.beg) and int(pf.total) == int(old.total)
            assert a_lists(pf.finish()) == a_lists(old.finish())
        assert totals.tolist() == [int(pf_e.total), int(pf_r.total)]
        # both lists in one launch (regcn_filter_fill2, what test() calls per timestamp) == the two single fills
        tot = totals.tolist()
        for one, two in zip((pf_e.finish(tot[0]), pf_r.finish(tot[1])), utils.filter_lists_finish2(pf_e, tot[0], pf_r, tot[1])):
            assert torch.equal(one.idx, two.idx) and torch.equal(one.end, two.end) and torch.equal(one.ptr, two.ptr)
            assert all(torch.equal(x, y) for x, y in zip(one.pairs(None), two.pairs(None)))


def a_lists(f):
    return f.lists()


def test_queries_prepare_batch_equals_per_snapshot_calls():
    """regcn_queries_prepare_batch (the test snapshots of a group of timestamps in three launches) against one
    regcn_queries_prepare per snapshot: inverse triples, list offsets, totals and the filled lists, incl. snapshots of one
    triple, of more queries than one key tile, and the maximum of 32 members."""
    R, _ = _ops()
    from regcn_b200 import utils
    rng = np.random.default_rng(5)
    for Ts, n, r in (((1,), 50, 4), ((37, 1, 5000, 12), 300, 7), (tuple(int(x) for x in rng.integers(1, 1600, 32)), 23033, 256)):
        snaps = [np.stack((rng.integers(0, n, T), rng.integers(0, r, T), rng.integers(0, n, T)), 1).astype(np.int64) for T in Ts]
        cat = torch.from_numpy(np.concatenate(snaps)).to(DEV)
        members, totals = utils.queries_prepare_batch(cat, list(Ts), r)
        tot = totals.tolist()
        for g, (snap, (all_t, pf_e, pf_r)) in enumerate(zip(snaps, members)):
            one_all, one_e, one_r, one_tot = utils.queries_prepare(torch.from_numpy(snap).to(DEV), r)
            assert torch.equal(all_t, one_all) and tot[g] == one_tot.tolist()
            assert torch.equal(pf_e.beg, one_e.beg) and torch.equal(pf_r.beg, one_r.beg)
            for mine, ref in zip(utils.filter_lists_finish2(pf_e, tot[g][0], pf_r, tot[g][1]),
                                 (one_e.finish(one_tot.tolist()[0]), one_r.finish(one_tot.tolist()[1]))):
                assert torch.equal(mine.idx, ref.idx) and torch.equal(mine.end, ref.end)
    with pytest.raises(Exception):
        utils.queries_prepare_batch(cat, [1] * 33, r)


# ----------------------------------------------------------------------------------------- tcgen05 GEMM
@pytest.mark.parametrize("impl,rtol", [("tc", 1e-4), ("tc1", 4e-3)])
@pytest.mark.parametrize("M,N,K,trans_b,bias,split_k", [
    (128, 16, 32, True, False, 1), (1, 200, 200, False, False, 1), (513, 200, 200, False, True, 1),
    (300, 600, 400, True, True, 1), (1000, 23033 // 8, 200, True, False, 1), (257, 200, 10000, True, True, 6),
    (64, 8, 12, False, False, 1), (130, 131, 204, True, True, 3), (2914, 400, 200, False, False, 1)])
def test_gemm_tcgen05_vs_fp64(impl, rtol, M, N, K, trans_b, bias, split_k):
    """tcgen05 kind::tf32 GEMM: 3xTF32 meets the fp32 parity gate, single-pass TF32 is reported at its own tolerance."""
    _, ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g)
    b = torch.randn((N, K) if trans_b else (K, N), generator=g)
    bv = torch.randn(N, generator=g) if bias else None
    ref = a.double() @ (b.double().t() if trans_b else b.double())
    if bias:
        ref = ref + bv.double()
    prev = ops.gemm_impl()
    ops.set_gemm_impl(impl)
    try:
        out = ops.gemm(a.to(DEV), b.to(DEV), trans_b=trans_b, bias=bv.to(DEV) if bias else None, split_k=split_k)
        torch.cuda.synchronize()
        tol = rtol * max(1.0, (K / 200.0) ** 0.5) * (4 if K > 2048 else 2)   # (unit-variance operands: the sum of |a||b| grows with K, 3xTF32 keeps ~2^-21 of it)
        err = (out.cpu().double() - ref).abs()
        scale = torch.clamp(ref.abs(), min=1.0) if impl == "tc" else torch.clamp(ref.abs(), min=float(K) ** 0.5)
        assert bool((err <= tol * scale).all()), float((err / scale).max())
        big = torch.zeros(M, N + 8, device=DEV)
        view = big[:, 4:4 + N]
        view.copy_(torch.ones(M, N))
        ops.gemm(a.to(DEV), b.to(DEV), trans_b=trans_b, out=view, accumulate=True, split_k=split_k)
        ref2 = a.double() @ (b.double().t() if trans_b else b.double()) + 1.0
        err2 = (view.cpu().double() - ref2).abs()
        scale2 = torch.clamp(ref2.abs(), min=1.0) if impl == "tc" else torch.clamp(ref2.abs(), min=float(K) ** 0.5)
        assert bool((err2 <= tol * scale2).all()), float((err2 / scale2).max())
        assert float(big[:, :4].abs().sum()) == 0.0 and float(big[:, 4 + N:].abs().sum()) == 0.0
    finally:
        ops.set_gemm_impl(prev)


def test_gemm_tcgen05_static_weight_cache():
    """A cached (transposed + split) weight is refreshed when the parameter changes in place."""
    _, ops = _ops()
    prev = ops.gemm_impl()
    ops.set_gemm_impl("tc")
    try:
        w = torch.nn.Parameter(torch.randn(200, 200, device=DEV))
        a = torch.randn(77, 200, device=DEV)
        o1 = ops.gemm(a, w, b_key=(w, "w"))
        with torch.no_grad():
            w.mul_(2.0)
        o2 = ops.gemm(a, w, b_key=(w, "w"))
        assert torch.allclose(o2, 2 * o1, rtol=1e-5, atol=1e-5)
    finally:
        ops.set_gemm_impl(prev)


# ----------------------------------------------------------------------------------------- bf16 scoring mode
@pytest.mark.parametrize("hyp", [False, True])
def test_bf16_scoring_counts_bracketed_by_fp64_on_rounded_operands(hyp):
    """bf16 scoring mode (tcgen05 kind::f16): the counts must be the exact 'beats the target' counts of SOME score
    matrix within 2e-5 (relative to the row scale) of the fp64 scores of the bf16-ROUNDED operands -- i.e. the only
    freedom is the fp32 accumulation order.  Stated separately from the fp32-parity mode (north_star)."""
    R, ops = _ops()
    rng = np.random.default_rng(21)
    B, N, d = 300, 5000, 200
    q = torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32) * 0.3).to(DEV)
    e = torch.from_numpy(rng.standard_normal((N, d)).astype(np.float32) * 0.3).to(DEV)
    target = torch.from_numpy(rng.integers(0, N, B).astype(np.int32)).to(DEV)
    ptr_t = torch.arange(B + 1, device=DEV, dtype=torch.int32) * 0
    idx_t = torch.zeros(1, device=DEV, dtype=torch.int32)
    pa = torch.arange(B, device=DEV, dtype=torch.int32)
    hyp_t = None
    if hyp:
        c = 0.01
        sm = torch.tensor([1.3, 0.7], device=DEV)
        hyp_t = (c, ops.row_sumsq(q), ops.row_sumsq(e), sm)
    prev = ops.score_dtype()
    ops.set_score_dtype("bf16")
    try:
        raw, filt, ts = ops.fused_rank_counts(q, e, target, ptr_t, idx_t, pa, target, hyp=hyp_t)
    finally:
        ops.set_score_dtype(prev)
    assert torch.equal(raw, filt)
    qb, eb = q.bfloat16().double().cpu(), e.bfloat16().double().cpu()
    dot = qb @ eb.T
    if hyp:
        x2 = (q.double().cpu() ** 2).sum(1, keepdim=True)
        y2 = (e.double().cpu() ** 2).sum(1).unsqueeze(0)
        xy = -dot
        a = 1 + 2 * c * xy + c * y2
        b = 1 - c * x2
        num = (a * a * x2 + 2 * a * b * xy + b * b * y2).clamp(min=0)
        den = 1 + 2 * c * xy + c * c * x2 * y2 + 1e-6
        n = torch.minimum(num.sqrt() / den.abs(), torch.tensor(1 / np.sqrt(c) - 2e-6, dtype=torch.float64))
        S = 1.3 * (0.7 - n * n)
    else:
        S = dot
    t = target.long().cpu()
    st = S[torch.arange(B), t].unsqueeze(1)
    eps = 2e-5 * S.abs().max(1, keepdim=True).values.clamp(min=1.0)
    lo = (S > st + eps).sum(1)
    hi = (S > st - eps).sum(1) - 1                         # the target itself is never counted
    r = raw.long().cpu()
    assert bool(((r >= lo) & (r <= hi)).all()), int(((r < lo) | (r > hi)).sum())
    ok, worst = close(ts.cpu().numpy(), st.squeeze(1).numpy(), rtol=2e-5, atol_scale=float(S.abs().max()))
    assert ok, worst


@pytest.mark.parametrize("a_mn,b_mn", [(1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K,split_k", [(200, 400, 1001, 1), (200, 200, 23033, 16), (37, 600, 512, 1), (130, 10000, 2914, 4)])
def test_gemm_mn_major_operands(a_mn, b_mn, M, N, K, split_k):
    """regcn_gemm_tf32_mn: operands in their natural row-major layout as MN-major tcgen05 tiles (dW = x^T dy, y = x W)
    against fp64, 3xTF32 (1e-4 relative to max(1,|ref|) scaled by sqrt(K))."""
    from regcn_b200 import _lib, ops
    _lib.require_device()
    rng = np.random.default_rng(M + N + K)
    pad = lambda n: (n + 3) // 4 * 4
    A = rng.standard_normal((K, M) if a_mn else (M, K)).astype(np.float32)
    B = rng.standard_normal((K, N) if b_mn else (N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)

    def dev_padded(x):
        r, c = x.shape
        t = torch.zeros((r, pad(c)), device=DEV)
        t[:, :c] = torch.from_numpy(x).to(DEV)
        return t

    Ad, Bd = dev_padded(A), dev_padded(B)
    a_hi, a_lo = ops.split_tf32(Ad)
    b_hi, b_lo = ops.split_tf32(Bd)
    ldc = pad(N)
    C = torch.full((M, ldc), 7.0, device=DEV)
    ws_bytes = _lib.load().regcn_gemm_tf32_workspace_bytes(M, N, split_k)
    ws = torch.empty(max(ws_bytes // 4, 1), device=DEV)
    _lib.call("regcn_gemm_tf32_mn", a_hi.data_ptr(), a_lo.data_ptr(), Ad.stride(0), b_hi.data_ptr(), b_lo.data_ptr(),
              Bd.stride(0), C.data_ptr(), ldc, M, N, K, a_mn, b_mn, torch.from_numpy(bias).to(DEV).data_ptr(), 0, 3, split_k,
              ws.data_ptr(), ws_bytes)
    ref = (A.astype(np.float64).T if a_mn else A.astype(np.float64)) @ (B.astype(np.float64) if b_mn else B.astype(np.float64).T) + bias
    got = C[:, :N].cpu().numpy()
    # the tensor core accumulates aligned products with ~20 bits below the largest term of a k-block, so the 3xTF32
    # sum sits ~2^-19 * sum_k |a||b| from fp64 (measured 4e-4 at K=512, 1e-3 at K=1001 for unit-variance operands) --
    # the same as the K-major kernel, which the second assertion pins exactly
    assert np.max(np.abs(got - ref)) <= 2e-6 * K + 1e-5, np.max(np.abs(got - ref))
    # ... and it is the SAME arithmetic as the K-major kernel on explicitly transposed copies: identical k-blocks,
    # identical products, identical accumulation order -> bit-identical results
    At = dev_padded(np.ascontiguousarray(A.T) if a_mn else A)
    Bt = dev_padded(np.ascontiguousarray(B.T) if b_mn else B)
    at_hi, at_lo = ops.split_tf32(At)
    bt_hi, bt_lo = ops.split_tf32(Bt)
    C2 = torch.full((M, ldc), 7.0, device=DEV)
    _lib.call("regcn_gemm_tf32", at_hi.data_ptr(), at_lo.data_ptr(), At.stride(0), bt_hi.data_ptr(), bt_lo.data_ptr(),
              Bt.stride(0), C2.data_ptr(), ldc, M, N, pad(K), torch.from_numpy(bias).to(DEV).data_ptr(), 0, 3, split_k,
              ws.data_ptr(), ws_bytes)
    same = torch.equal(C[:, :N], C2[:, :N])
    assert same or float((C[:, :N] - C2[:, :N]).abs().max()) <= 1e-6 * max(1.0, float(C2[:, :N].abs().max())), \
        float((C[:, :N] - C2[:, :N]).abs().max())


# ----------------------------------------------------------------------------------------- fp32-A GEMM (on-chip TF32 split)
@pytest.mark.parametrize("M,N,k0,k1,gather,split_k", [
    (300, 200, 200, 0, False, 1), (23033, 400, 200, 0, False, 1), (1560, 200, 200, 200, True, 1),
    (129, 64, 36, 8, True, 1), (2914, 200, 1000, 0, False, 3), (5, 16, 4, 0, False, 1)])
def test_gemm_fp32_a_bit_identical_to_presplit(M, N, k0, k1, gather, split_k):
    """regcn_gemm_tf32_a32 (A split into TF32 hi/lo inside shared memory by converter warps, optional row gather, two K
    segments) must equal regcn_gemm_tf32 on the pre-split, pre-gathered, concatenated operand BIT FOR BIT."""
    R, ops = _ops()
    from regcn_b200 import _lib
    g = torch.Generator(device=DEV)
    g.manual_seed(M * 7 + N)
    src_rows = M + 77 if gather else M
    a0 = torch.randn(src_rows, k0 + 4, device=DEV, generator=g)[:, :k0]          # lda > k
    a1 = torch.randn(src_rows, max(k1, 4), device=DEV, generator=g)
    rows0 = torch.randperm(src_rows, device=DEV, generator=g)[:M].to(torch.int32) if gather else None
    rows1 = torch.randperm(src_rows, device=DEV, generator=g)[:M].to(torch.int32) if gather and k1 else None
    K = k0 + k1
    w = torch.randn(N, K, device=DEV, generator=g)
    bias = torch.randn(N, device=DEV, generator=g)
    parts = [a0[rows0.long()] if gather else a0]
    if k1:
        parts.append((a1[rows1.long()] if rows1 is not None else a1[:M])[:, :k1])
    a_cat = torch.cat(parts, dim=1).contiguous()
    ah, al = ops.split_tf32(a_cat) if K % 4 == 0 else (None, None)
    wh, wl = ops.split_tf32(w)
    ref = torch.empty(M, N, device=DEV)
    ws = torch.empty(max(1, split_k * M * N), device=DEV)
    _lib.call("regcn_gemm_tf32", ah.data_ptr(), al.data_ptr(), K, wh.data_ptr(), wl.data_ptr(), K, ref.data_ptr(), N, M, N, K,
              bias.data_ptr(), 0, 3, split_k, ws.data_ptr(), ws.numel() * 4)
    out = torch.full((M, N), float("nan"), device=DEV)
    _lib.call("regcn_gemm_tf32_a32", a0.data_ptr(), a0.stride(0), k0, None if rows0 is None else rows0.data_ptr(),
              a1.data_ptr() if k1 else None, a1.stride(0), k1, None if rows1 is None else rows1.data_ptr(), wh.data_ptr(),
              wl.data_ptr(), K, out.data_ptr(), N, M, N, bias.data_ptr(), 0, 3, split_k, ws.data_ptr(), ws.numel() * 4, None, 0)
    torch.cuda.synchronize()
    if k0 % 32 == 0 or k1 == 0:
        assert torch.equal(out, ref), f"max |diff| {(out - ref).abs().max().item()}"
    else:
        # two segments with a ragged first segment: k-blocks are padded per segment, so products are accumulated in a
        # different k-block grouping than on the concatenated operand -- same values to fp32 rounding
        ok, worst = close(out.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5)
        assert ok, worst


def test_gemm_layer_fp32_a_matches_presplit_layer():
    """Layer epilogue (rrelu, gate columns, row scatter / skip, fused normalise + time gate) through the fp32-A path equals
    the pre-split path bit for bit."""
    R, ops = _ops()
    from regcn_b200 import _lib
    g = torch.Generator(device=DEV)
    g.manual_seed(5)
    N, d = 3000, 200
    x = torch.randn(N, d, device=DEV, generator=g) * 0.3
    w = torch.randn(2 * d, d, device=DEV, generator=g) * 0.1
    xh, xl = ops.split_tf32(x)
    wh, wl = ops.split_tf32(w)
    skip = torch.full((N,), -1, device=DEV, dtype=torch.int32)
    skip[::7] = 1
    outs = []
    for a32 in (False, True):
        o_raw = torch.zeros(N, d, device=DEV)
        gate = torch.zeros(N, d, device=DEV)
        if a32:
            _lib.call("regcn_gemm_tf32_layer_a32", x.data_ptr(), d, d, None, None, 0, 0, None, wh.data_ptr(), wl.data_ptr(), d,
                      N, 2 * d, d, o_raw.data_ptr(), None, None, gate.data_ptr(), d, None, skip.data_ptr(), None, 0, None,
                      None, 0)
        else:
            _lib.call("regcn_gemm_tf32_layer", xh.data_ptr(), xl.data_ptr(), d, wh.data_ptr(), wl.data_ptr(), d, N, 2 * d, d,
                      d, o_raw.data_ptr(), None, None, gate.data_ptr(), d, None, skip.data_ptr(), None, 0, None, None, 0)
        # last layer: normalise + time gate on the layer-0 gate columns
        h_new = torch.zeros(N, d, device=DEV)
        b = torch.randn(d, device=DEV, generator=torch.Generator(device=DEV).manual_seed(9)) * 0.1
        w1 = torch.randn(d, d, device=DEV, generator=torch.Generator(device=DEV).manual_seed(11)) * 0.1
        w1h, w1l = ops.split_tf32(w1)
        if a32:
            _lib.call("regcn_gemm_tf32_layer_a32", o_raw.data_ptr(), d, d, None, None, 0, 0, None, w1h.data_ptr(),
                      w1l.data_ptr(), d, N, d, d, h_new.data_ptr(), None, None, None, 0, None, skip.data_ptr(),
                      gate.data_ptr(), d, b.data_ptr(), x.data_ptr(), 1)
        else:
            oh, ol = ops.split_tf32(o_raw)
            _lib.call("regcn_gemm_tf32_layer", oh.data_ptr(), ol.data_ptr(), d, w1h.data_ptr(), w1l.data_ptr(), d, N, d, d, d,
                      h_new.data_ptr(), None, None, None, 0, None, skip.data_ptr(), gate.data_ptr(), d, b.data_ptr(),
                      x.data_ptr(), 1)
        torch.cuda.synchronize()
        outs.append((o_raw, gate, h_new))
    for a, b_ in zip(*outs):
        assert torch.equal(a, b_), (a - b_).abs().max().item()
    # and against torch in fp64
    ref1 = torch.nn.functional.rrelu((x.double() @ w[:d].double().t()), training=False)
    act = skip.cpu().numpy() < 0
    ok, worst = close(outs[1][0].cpu().numpy()[act], ref1.float().cpu().numpy()[act], rtol=1e-4)
    assert ok, worst


def test_split_tf32_matches_rna_restatement():
    """regcn_split_tf32 (bit-pattern rna) == the frexp/floor restatement of cvt.rna.tf32.f32 on normal values, ties and
    signs included; hi and lo have their 13 low mantissa bits clear and hi + lo reproduces x to 2^-21."""
    R, ops = _ops()
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.standard_normal(40000).astype(np.float32) * s for s in (1e-20, 1e-3, 1.0, 3e4, 1e20)])
    ties = (np.arange(1, 4097, dtype=np.uint32) << 13 | 0x1000 | 0x3f800000).view(np.float32)   # exactly half-way cases
    x = np.concatenate([x, ties, -ties, np.array([0.0, -0.0, 1.0, -1.0, 0.1, 1e-30], dtype=np.float32)]).astype(np.float32)
    x = np.resize(x, (x.size + 3) // 4 * 4)
    hi, lo = ops.split_tf32(_t(x))
    hi, lo = hi.cpu().numpy(), lo.cpu().numpy()
    r_hi, r_lo = restate.split_tf32(x)
    assert np.array_equal(hi, r_hi)
    assert np.array_equal(lo, r_lo)
    assert not (hi.view(np.uint32) & 0x1fff).any() and not (lo.view(np.uint32) & 0x1fff).any()
    assert np.all(np.abs(hi.astype(np.float64) + lo - x) <= 2.0 ** -21 * np.abs(x))


@pytest.mark.parametrize("B,N,scale,margin,rmax", [(300, 5000, 1.0, 1.0, 3.0), (129, 2049, 0.37, 4.0, 9.9),
                                                  (64, 777, 2.5, 0.0, 0.3), (8192, 20000, 1.0, 1.0, 3.0)])
def test_hyp_count_threshold_polynomial_equals_exact_scores(B, N, scale, margin, rmax):
    """The division-free polynomial threshold test of the hyperbolic counting epilogue (regcn_score_count_poly(1)) gives
    exactly the counts of the IEEE score evaluated for every candidate (regcn_score_count_poly(0)) -- on random points of
    the ball, on exact ties (duplicated target rows), near-ties (rows one ulp away) and targets at the projection clamp."""
    R, ops = _ops()
    from regcn_b200 import _lib
    lib = _lib.load()
    d, c = 200, 0.01
    g = torch.Generator(device=DEV)
    g.manual_seed(B + N)

    def ball(n):
        x = torch.randn(n, d, device=DEV, generator=g)
        r = torch.rand(n, 1, device=DEV, generator=g) * rmax
        return x / x.norm(dim=1, keepdim=True) * r

    q, e = ball(B), ball(N)
    target = torch.randint(0, N, (B,), device=DEV, dtype=torch.int32, generator=g)
    # exact ties and near ties of the target row; a few candidates at / beyond the ball's boundary
    for k in range(0, B, 3):
        t = int(target[k])
        e[(t + 17) % N] = e[t]
        e[(t + 31) % N] = e[t] * (1.0 + 1.2e-7)
        e[(t + 47) % N] = torch.nextafter(e[t], torch.full_like(e[t], 100.0))
    e[5] = e[5] / e[5].norm() * 9.999995
    e[6] = e[6] / e[6].norm() * 10.5
    q[1] = q[1] / q[1].norm() * 9.9999
    target[2] = 5
    target[3] = 6
    qh, ql = ops.split_tf32(q.contiguous())
    eh, el = ops.split_tf32(e.contiguous())
    x2, y2 = ops.row_sumsq(q), ops.row_sumsq(e)
    sm = torch.tensor([scale, margin], device=DEV)
    # target scores through the pair kernel (the arithmetic of the counted scores), like ops.fused_rank_counts
    idx_a = torch.arange(B, device=DEV, dtype=torch.int32)
    ah, al = torch.empty(B, d, device=DEV), torch.empty(B, d, device=DEV)
    bh, bl = torch.empty(B, d, device=DEV), torch.empty(B, d, device=DEV)
    _lib.call("regcn_gather_rows2", qh.data_ptr(), ql.data_ptr(), idx_a.data_ptr(), B, d, ah.data_ptr(), al.data_ptr())
    _lib.call("regcn_gather_rows2", eh.data_ptr(), el.data_ptr(), target.data_ptr(), B, d, bh.data_ptr(), bl.data_ptr())
    y2p = y2[target.long()].contiguous()
    ts = torch.empty(B, device=DEV)
    _lib.call("regcn_pair_scores_tf32", ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), B, d, 1, x2.data_ptr(),
              y2p.data_ptr(), None, c, sm.data_ptr(), None, ts.data_ptr(), 3)
    counts = []
    try:
        for on in (0, 1):
            lib.regcn_score_count_poly(on)
            raw = torch.zeros(B, device=DEV, dtype=torch.int32)
            _lib.call("regcn_score_count_tf32", qh.data_ptr(), ql.data_ptr(), eh.data_ptr(), el.data_ptr(), B, N, d,
                      ts.data_ptr(), target.data_ptr(), raw.data_ptr(), 0, 1, x2.data_ptr(), y2.data_ptr(), None, c,
                      sm.data_ptr(), None, 3)
            torch.cuda.synchronize()
            counts.append(raw.cpu().numpy())
    finally:
        lib.regcn_score_count_poly(1)
    assert np.array_equal(counts[0], counts[1]), np.nonzero(counts[0] != counts[1])[0][:10]
    assert counts[0].max() > 0
