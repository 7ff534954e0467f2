"""Evaluation step of the hot path as one call: evolve the history, score every entity, rank raw + filtered.

This is the loop body of the reference's test() (src/main.py:67-74, hyperbolic_main.py:100-113) for entity
prediction, expressed on device-resident inputs so that nothing but kernels runs inside it.
"""
import torch

from . import ops, utils


def _scoring_operands(model, emb, r_emb, all_triples, normalize=False):
    """(q, cand, hyp, col_bias) of the entity decoder: everything the scoring GEMM needs except the GEMM itself.
    normalize (ConvTransE only): emb is the evolved table before the predict-time F.normalize, folded into the tanh pass."""
    dec = model.decoder_ob
    name = type(dec).__name__
    if name == "ConvTransE":
        e_all, q = dec.query(emb, r_emb, all_triples, normalize=normalize)
        return q, e_all, None, None
    assert not normalize
    if name == "HyperbolicConvTransE":
        et = ops.row_map(emb, ops.ROW_LEAKY_TANH_LOG0, c=dec.c)
        q = dec._tower(et, r_emb.contiguous(), all_triples, 0, 1, always_bn2=False)
        return q, et, None, dec.b.detach()
    if name in ("HyperbolicRotH", "HyperbolicMuRP", "HyperbolicAttH"):
        q, qss = dec.query(emb, r_emb, all_triples)
        cand = emb.contiguous()
        # entity_bias[subject] shifts a whole row and cannot change a rank: only the candidate bias is needed here
        return q, cand, dec.hyp_operands(qss, cand, all_triples), dec._entity_bias(all_triples)[0]
    raise NotImplementedError(name)


def _tower_table(dec):
    """Host array of the 13 device pointers regcn_convtrans_decode_rank takes per decoder (include/regcn_b200.h), rebuilt
    when any of the module's tensors changes (version / storage)."""
    import numpy as np
    from .decoder import _fold_bn
    srcs = [dec.conv1.weight, dec.conv1.bias, dec.fc.weight, dec.fc.bias]
    for bn in (dec.bn0, dec.bn1, dec.bn2):
        srcs += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
    stamp = tuple((t._version, t.data_ptr()) for t in srcs)
    hit = dec.__dict__.get("_regcn_tower_tab")
    if hit is not None and hit[0] == stamp:
        return hit[1]
    with torch.no_grad():
        s0, b0 = _fold_bn(dec.bn0)
        s1, b1 = _fold_bn(dec.bn1)
        s2, b2 = _fold_bn(dec.bn2)
        cw = dec.conv1.weight.detach().contiguous()
        cb = dec.conv1.bias.detach().contiguous()
        fw_hi, fw_lo = ops.split_tf32(dec.fc.weight.detach().contiguous())
        fb = dec.fc.bias.detach().contiguous()
        C_, _, ksz_ = cw.shape
        d_ = dec.fc.in_features // C_
        if ops.convtrans_fc_ok(d_, cw, dec.fc.out_features):
            fz_hi, fz_lo = ops.convtrans_fc_weight(dec.fc.weight, C_, d_)     # reduction order of the fused tower
        else:
            fz_hi, fz_lo = fw_hi, fw_lo
    keep = [s0, b0, cw, cb, s1, b1, fw_hi, fw_lo, fb, s2, b2, fz_hi, fz_lo]
    tab = np.array([t.data_ptr() for t in keep], dtype=np.uint64)
    dec.__dict__["_regcn_tower_tab"] = (stamp, (tab, keep))
    return tab, keep


def convtrans_decode_rank(model, emb, r_emb, all_t, f_ent, f_rel):
    """The decode + rank half of one evaluated timestamp (src/main.py:71-74 after the evolution) in ONE C-ABI call
    (`regcn_convtrans_decode_rank`).  emb: the evolved entity table BEFORE the predict-time F.normalize.  Returns the
    packed int32 ranks [rank | filter_rank | rank_rel | filter_rank_rel] (4B,) on the device."""
    from . import _lib
    dec, rdec = model.decoder_ob, model.rdecoder
    emb, r_emb = emb.contiguous(), r_emb.contiguous()
    N, d = emb.shape
    R2, B = r_emb.shape[0], all_t.shape[0]
    C, _, ksz = dec.conv1.weight.shape
    pa, pe = f_ent.pairs(None if f_ent._pairs is not None else all_t[:, 2].to(torch.int32).contiguous())
    P = int(pa.shape[0])
    tab_e, _keep_e = _tower_table(dec)
    tab_r, _keep_r = _tower_table(rdec)
    need = _lib.load().regcn_convtrans_decode_rank_workspace_bytes(N, R2, d, B, C, P)
    ws = model.__dict__.get("_regcn_decode_ws")
    if ws is None or ws.numel() < need or ws.device != emb.device:
        ws = torch.empty(int(need * 1.25), device=emb.device, dtype=torch.uint8)
        model.__dict__["_regcn_decode_ws"] = ws
    packed = torch.empty(4 * B, device=emb.device, dtype=torch.int32)
    _lib.call("regcn_convtrans_decode_rank", emb.data_ptr(), r_emb.data_ptr(), all_t.data_ptr(), tab_e.ctypes.data,
              tab_r.ctypes.data, f_ent.ptr.data_ptr(), f_ent.idx.data_ptr(),
              None if f_ent.end is None else f_ent.end.data_ptr(), pa.data_ptr(), pe.data_ptr(), P, f_rel.ptr.data_ptr(),
              f_rel.idx.data_ptr(), None if f_rel.end is None else f_rel.end.data_ptr(), N, R2, d, B, C, ksz,
              int(bool(model.layer_norm)), packed.data_ptr(), ws.data_ptr(), ws.numel())
    return packed


@torch.no_grad()
def _score_rank(model, emb, r_emb, all_triples, filter_csr, fused, shard, mark):
    """Scoring + ranking half of one evaluated timestamp; emb is the evolved entity table before the predict-time
    normalisation (src/rrgcn.py:186-194)."""
    mark("score", 0)
    fold_norm = bool(model.layer_norm) and fused and type(model.decoder_ob).__name__ == "ConvTransE"
    if model.layer_norm and not fold_norm:
        if hasattr(model, "_c_float"):
            emb = ops.row_map(emb, ops.ROW_TANGENT_NORMALIZE, c=model._c_float)
        else:
            emb = ops.row_map(emb, ops.ROW_NORMALIZE)
    if fused:
        q, cand, hyp, col_bias = _scoring_operands(model, emb, r_emb, all_triples, normalize=fold_norm)
        target = all_triples[:, 2].to(torch.int32).contiguous()
        pa, pe = filter_csr.pairs(target)
        mark("score", 1)
        mark("rank", 0)
        raw, filt, _ = ops.fused_rank_counts(q, cand, target, filter_csr.ptr, filter_csr.idx, pa, pe, hyp=hyp,
                                             col_bias=col_bias, shard=shard, filt_end=filter_csr.end)
        if shard is not None:
            mark("rank", 1)
            return raw, filt
        rank, frank = ops.counts_to_ranks(raw, filt)
        mark("rank", 1)
        return rank, frank
    score = model.decoder_ob.forward(emb, r_emb, all_triples, mode="test")
    mark("score", 1)
    mark("rank", 0)
    raw, filt, _ = ops.rank_dense(score, all_triples, 2, filter_csr.ptr if filter_csr is not None else None,
                                  filter_csr.idx if filter_csr is not None else None,
                                  filt_end=filter_csr.end if filter_csr is not None else None)
    rank, frank = ops.counts_to_ranks(raw, filt)
    mark("rank", 1)
    return rank, frank


@torch.no_grad()
def evaluate_snapshot(model, glist, all_triples, filter_csr, timers=None, fused=None, shard=None):
    """Returns (rank, filter_rank) int64 (B,) for the 2*T_q queries `all_triples` (forward + inverse).

    fused (default: on with the tensor-core GEMM): rank through the counting epilogue of the scoring GEMM, never
    writing the (B,N) score matrix; otherwise the dense predict()-style path.  shard=(lo,hi) restricts the counted
    candidates (entity-sharded scoring: the caller all-reduces the counts).
    timers: optional dict of name -> (start_event, end_event) pairs recorded on the current stream."""
    def mark(name, which):
        if timers is not None:
            timers[name][which].record()

    if fused is None:
        fused = ops.gemm_impl() in ("tc", "tc1") and filter_csr is not None
    mark("evolve", 0)
    evolve_embs, _, r_emb, _, _ = model.forward(glist, None, True)
    mark("evolve", 1)
    return _score_rank(model, evolve_embs[-1], r_emb, all_triples, filter_csr, fused, shard, mark)


BATCH_ROWS = 192 * 1024     # entity rows one batched recurrence aims for (timestamps_per_batch)


def timestamps_per_batch(model, num_nodes, static_graph=None):
    """How many consecutive test timestamps test() / evaluate_batch evolve at once.  At TKG sizes (7-23 k entities) the
    kernels of one timestamp's recurrence are latency-bound; G independent windows as one block-diagonal graph give every
    launch G times the rows (measured at the ICEWS18 shape: 0.85 ms per timestamp alone, 0.50 ms in batches of 8).
    REGCN_TEST_BATCH=<n> overrides; 1 = one timestamp per recurrence."""
    import os
    ok = hasattr(model, "forward_batch") and model.batch_ok() and not (static_graph is not None and model.use_static)
    if not ok:
        return 1
    env = os.environ.get("REGCN_TEST_BATCH")
    if env:
        return max(1, min(32, int(env)))
    if hasattr(model, "_forward_engine_shared") and os.environ.get("REGCN_SHARED_ROWS", "1") != "0":
        # shared-trajectory engine (regcn_regcn_evolve_shared): the all-entity products run over the rows touched so far,
        # not G N rows, and the shared rows are amortised over the windows (ICEWS18 shape: 261 / 225 / 208 us per
        # timestamp at 8 / 12 / 16 windows)
        return max(1, min(32, 4 * BATCH_ROWS // max(1, int(num_nodes))))
    # full block-diagonal recurrence (the hyperbolic model; RE-GCN outside the shared engine's preconditions): measured at
    # the ICEWS14s shape (lgcn / hyperbolic_uvrgcn, profiles/time_batched_forward_hyp.py) 732 / 669 us per timestamp alone,
    # 301 / 321 at 8 windows, 241 / 266 at 16, 219 / 243 at 24
    return max(1, min(24, BATCH_ROWS // max(1, int(num_nodes))))


def group_sizes(K, G, ramp="8"):
    """How test() cuts K consecutive test timestamps into groups that share one recurrence: the short first groups named
    by `ramp` (comma-separated sizes, each used only while it is below G), then the rest in equal shares of at most G."""
    if G <= 1:
        return [1] * K
    sizes, left = [], K
    for cap in (int(x) for x in str(ramp).split(",") if x.strip()):
        if left > 0 and 0 < cap < G:
            sizes.append(min(cap, left))
            left -= sizes[-1]
    if left > 0:
        n_rest = -(-left // G)
        sizes += [left // n_rest + (1 if i < left % n_rest else 0) for i in range(n_rest)]
    return sizes


@torch.no_grad()
def evaluate_batch(model, windows, all_triples_list, filter_list, timers=None, fused=None):
    """evaluate_snapshot for G test timestamps whose history windows are evolved together (model.forward_batch): returns
    [(rank, filter_rank)] in the order given, each pair identical to evaluate_snapshot(model, windows[g], ...).
    timers: 'evolve' brackets the batched recurrence, 'score' everything after it."""
    def mark(name, which):
        if timers is not None and name in timers:
            timers[name][which].record()

    def no_mark(name, which):
        pass

    mark("evolve", 0)
    states = model.forward_batch(windows)
    mark("evolve", 1)
    mark("score", 0)
    out = []
    for (emb, r_emb), all_t, f in zip(states, all_triples_list, filter_list):
        fz = (ops.gemm_impl() in ("tc", "tc1") and f is not None) if fused is None else fused
        out.append(_score_rank(model, emb, r_emb, all_t, f, fz, None, no_mark))
    mark("score", 1)
    return out


QUERY_SHARD_MIN = 16384      # smallest query batch whose tower is cut across ranks (see score_rank_sharded)


@torch.no_grad()
def score_rank_sharded(model, emb, r_emb, all_triples, filter_csr, group=None):
    """Entity-partitioned scoring + rank merge: this rank counts over its contiguous slice of the
    (replicated, already evolved) entity table with the fused kernel; ONE all_reduce(SUM) of the (2,B) int32 counts
    merges the shards.  Target and filter-entry scores come from the pair pass on the replicated table, so no score
    exchange is needed.  Returns (rank, filter_rank), identical on every rank."""
    from . import dist as rdist
    r, ws = rdist.world(group)
    B = all_triples.shape[0]
    if ws > 1 and type(model.decoder_ob).__name__ == "ConvTransE" and B >= QUERY_SHARD_MIN:
        # query-sharded tower: this rank builds the queries of its B/G slice (conv features + split-K FC), one all_gather
        # of the (B,d) matrix over NVLink.  Only for large query batches: measured on 2 B200s at the ICEWS18 size
        # (B = 2914) the FC of a half batch takes as long as the whole one (73 vs 78 us: 313 k-blocks of latency, the
        # split-K factor is pinned to the whole batch so that a row's value does not depend on the cut) and the extra
        # collective costs more than the 30 us the feature kernel saves (profiles/README.md, round 2)
        b0, b1 = rdist.shard_bounds(B, r, ws)
        cand, q_loc = model.decoder_ob.query(emb, r_emb, all_triples[b0:b1].contiguous(), batch_total=B)
        q = rdist.gather_rows(q_loc, B, group).contiguous()
        hyp = col_bias = None
    else:
        q, cand, hyp, col_bias = _scoring_operands(model, emb, r_emb, all_triples)
    lo, hi = rdist.shard_bounds(cand.shape[0], r, ws)
    target = all_triples[:, 2].to(torch.int32).contiguous()
    pa, pe = filter_csr.pairs(target)
    raw, filt, _ = ops.fused_rank_counts(q, cand, target, filter_csr.ptr, filter_csr.idx, pa, pe, hyp=hyp,
                                         col_bias=col_bias, shard=(lo, hi), filt_end=filter_csr.end)
    return rdist.merge_counts(raw, filt, group)


@torch.no_grad()
def evaluate_from_host(model, history_host, test_host, num_nodes, num_rels, device):
    """End-to-end evaluation of one test timestamp from HOST buffers (pinned int64 triples), the loop body of the
    reference's test() (src/main.py:67-74): H2D copies, device edge-index build for every history snapshot (one
    host sync for all of them), evolution, entity ranks through the fused scoring kernel, relation ranks through
    ConvTransR / RotHRel + the dense rank kernel, time-aware filtering for both, and the D2H read-back of MRRs
    and rank vectors.  Returns ((filter_mrr, mrr, filter_mrr_rel, mrr_rel), rank_host, filter_rank_host)."""
    from .graph import build_sub_graphs, finish_sub_graphs, pending_counts
    # every index build is enqueued first (history graphs: one batched launch; filter lists: count pass), then ONE
    # device->host read returns all the sizes the host needs (per-snapshot counters, filter slot totals)
    glist = build_sub_graphs(num_nodes, num_rels, history_host, device, sync=False)
    test = test_host.to(device, non_blocking=True)
    inv = test.flip(1)
    inv[:, 1] = inv[:, 1] + num_rels
    all_t = torch.cat((test, inv)).contiguous()
    if 0 < all_t.shape[0] <= 32768:
        pf_ent, pf_rel = utils.filter_lists_begin(all_t, 0), utils.filter_lists_begin(all_t, 1)
        sizes = torch.cat((pending_counts(glist).flatten(), pf_ent.total, pf_rel.total)).tolist()
        L = len(glist)
        finish_sub_graphs(glist, [sizes[8 * i:8 * i + 8] for i in range(L)])
        f_ent, f_rel = pf_ent.finish(sizes[8 * L]), pf_rel.finish(sizes[8 * L + 1])
    else:
        finish_sub_graphs(glist)
        f_ent = utils.filter_csr_from_snapshot(all_t, 2 * num_rels, 0, num_answers=num_nodes)
        f_rel = utils.filter_csr_from_snapshot(all_t, num_nodes, 1, num_answers=2 * num_rels)
    evolve_embs, _, r_emb, _, _ = model.forward(glist, None, True)
    emb = evolve_embs[-1]
    if model.layer_norm:
        if hasattr(model, "_c_float"):
            emb = ops.row_map(emb, ops.ROW_TANGENT_NORMALIZE, c=model._c_float)
        else:
            emb = ops.row_map(emb, ops.ROW_NORMALIZE)
    if ops.gemm_impl() in ("tc", "tc1"):
        q, cand, hyp, col_bias = _scoring_operands(model, emb, r_emb, all_t)
        target = all_t[:, 2].to(torch.int32).contiguous()
        pa, pe = f_ent.pairs(target)
        raw, filt, _ = ops.fused_rank_counts(q, cand, target, f_ent.ptr, f_ent.idx, pa, pe, hyp=hyp, col_bias=col_bias,
                                             filt_end=f_ent.end)
        rank, frank = ops.counts_to_ranks(raw, filt)
    else:
        score = model.decoder_ob.forward(emb, r_emb, all_t, mode="test")
        _, _, rank, frank = utils.get_total_rank(all_t, score, None, 1000, rel_predict=0, filter_csr=f_ent)
    score_rel = model.rdecoder.forward(emb, r_emb, all_t, mode="test")
    raw_r, filt_r, _ = ops.rank_dense(score_rel, all_t, 1, f_rel.ptr, f_rel.idx, filt_end=f_rel.end)
    rank_r, frank_r = ops.counts_to_ranks(raw_r, filt_r)
    mrrs = torch.stack([torch.mean(1.0 / frank.float()), torch.mean(1.0 / rank.float()),
                        torch.mean(1.0 / frank_r.float()), torch.mean(1.0 / rank_r.float())])
    out = torch.cat((mrrs.double(), rank.double(), frank.double())).cpu()   # one D2H transfer (ranks < 2^53 exactly)
    B = rank.numel()
    return tuple(out[:4].tolist()), out[4:4 + B].long(), out[4 + B:].long()


class _Prepared:
    """Inputs of one evaluated timestamp whose device-side preparation is enqueued but whose sizes are still on the way
    to the host (pinned buffer + event)."""
    __slots__ = ("glist", "new_graphs", "all_t", "pf_ent", "pf_rel", "sizes_host", "event", "totals", "group")


def _prepare(cache, input_list, test_snap, num_rels, device, sizes_slot, prep_stream, main_stream, ahead=()):
    """Enqueue the device-side preparation of one timestamp on `prep_stream` (so that it runs next to, not behind, the
    previous timestamp's kernels) and start the copy of its sizes into `sizes_slot`, a pinned int32 buffer allocated once
    per test() call (page-locking memory per step would cost more than the step).  Every tensor created here is handed
    to `main_stream` later: record_stream keeps the caching allocator from recycling it while that stream still reads."""
    from .graph import pending_counts
    p = _Prepared()
    import contextlib
    with (torch.cuda.stream(prep_stream) if prep_stream is not main_stream else contextlib.nullcontext()):
        # `ahead`: snapshots that will enter the window during the next steps; their indices are built in the same batched
        # launch (one CTA per snapshot, all concurrent) as whatever is missing now
        glist, p.new_graphs = cache.ensure(list(input_list) + list(ahead))
        p.glist = glist[:len(input_list)]
        t = test_snap if isinstance(test_snap, torch.Tensor) else torch.from_numpy(test_snap)
        test = t.to(device, non_blocking=True)
        if 0 < test.shape[0] <= 16384:
            # inverse triples, both filter-list count passes and their scans in one C call
            p.all_t, p.pf_ent, p.pf_rel, totals = utils.queries_prepare(test, num_rels)
            sizes = torch.cat((totals, pending_counts(p.new_graphs).flatten())) if p.new_graphs else totals
        else:
            inv = test.flip(1)
            inv[:, 1] = inv[:, 1] + num_rels
            p.all_t = torch.cat((test, inv)).contiguous()
            p.pf_ent, p.pf_rel = utils.filter_lists_begin(p.all_t, 0), utils.filter_lists_begin(p.all_t, 1)
            parts = [p.pf_ent.total, p.pf_rel.total]
            if p.new_graphs:
                parts.append(pending_counts(p.new_graphs).flatten())
            sizes = torch.cat(parts)
        p.sizes_host = sizes_slot[:sizes.numel()]
        p.sizes_host.copy_(sizes, non_blocking=True)
        p.event = torch.cuda.Event()
        p.event.record(prep_stream)
        if prep_stream is not main_stream:
            for x in [p.all_t, p.pf_ent.beg, p.pf_rel.beg, p.pf_ent.total, p.pf_rel.total] + \
                     [g._arena for g in p.new_graphs] + [g.triples for g in p.new_graphs]:
                x.record_stream(main_stream)
    return p


class _GroupPrep:
    """What the members of a group prepared by ONE call share: the graphs built for it, the pinned copy of its sizes
    ([entity / relation slot totals per member | 8 counters per new graph]) and the event behind that copy."""
    __slots__ = ("new_graphs", "sizes_host", "event", "members", "done")

    def finish(self):
        from .graph import finish_sub_graphs
        if self.done:
            return
        self.event.synchronize()
        sizes = self.sizes_host.tolist()
        n = len(self.members)
        if self.new_graphs:
            finish_sub_graphs(self.new_graphs, [sizes[2 * n + 8 * i:2 * n + 8 * i + 8] for i in range(len(self.new_graphs))])
        for g, p in enumerate(self.members):
            p.totals = (sizes[2 * g], sizes[2 * g + 1])
        self.done = True


def _prepare_group(cache, windows, test_snaps, num_rels, device, sizes_slot, stage=None, on_device=None):
    """Device-side preparation of the n <= 32 timestamps of a group as ONE batch on the current stream: every snapshot
    missing from the cache (history windows incl. the test snapshots that slide into them) uploaded in one staged copy and
    indexed by one batched launch, the n test snapshots uploaded in one copy and prepared by `regcn_queries_prepare_batch`
    (inverse triples, both filter count passes, their scans: three launches for the group), ONE device -> host copy of all
    sizes.  Per timestamp this was 2 copies + 3 launches + a size copy, each separated by launch / copy-engine latency on
    the stream the decodes run on (~45 us of the ~0.8 ms a timestamp takes end to end) and ~0.13 ms of host time.
    Returns the members (`_Prepared`, in order) or None when a snapshot is empty / too large for the batched call."""
    from .graph import pending_counts
    n = len(test_snaps)
    tens = [t if isinstance(t, torch.Tensor) else torch.from_numpy(t) for t in test_snaps]
    Ts = [int(t.shape[0]) for t in tens]
    if n < 1 or n > 32 or any(T <= 0 or T > 16384 for T in Ts) or any(t.dtype != torch.int64 or t.dim() != 2 for t in tens):
        return None
    flat = [s for w in windows for s in w]
    if all(not t.is_cuda for t in tens):
        # `stage`: this group's slice of a pinned buffer allocated once per test() call (a pinned allocation per group
        # costs more than it saves whenever the caching host allocator has to page-lock a fresh block)
        if stage is None or stage.numel() != 3 * sum(Ts):
            stage = torch.empty(3 * sum(Ts), dtype=torch.int64, pin_memory=True)
        torch.cat([t.reshape(-1) for t in tens], out=stage)
        cat = torch.empty((stage.numel() + 65535) // 65536 * 65536, device=device, dtype=torch.int64)[:stage.numel()]
        cat.copy_(stage, non_blocking=True)
        cat = cat.view(-1, 3)
    else:
        cat = torch.cat([t.to(device, non_blocking=True) for t in tens])
    if on_device is not None:
        # `on_device` {id(snapshot): device triples}: a test snapshot slides into the history windows of the following
        # timestamps (src/main.py:98-100); its index is built from the copy made here, not from a second upload
        o = 0
        for snap, T in zip(test_snaps, Ts):
            on_device[id(snap)] = cat[o:o + T]
            o += T
    glist_all, new_graphs = cache.ensure(flat, on_device)
    prepared, totals = utils.queries_prepare_batch(cat, Ts, num_rels)
    sizes = torch.cat((totals.flatten(), pending_counts(new_graphs).flatten())) if new_graphs else totals.flatten()
    grp = _GroupPrep()
    grp.new_graphs, grp.done = new_graphs, False
    grp.sizes_host = sizes_slot[:sizes.numel()]
    grp.sizes_host.copy_(sizes, non_blocking=True)
    grp.event = torch.cuda.Event()
    grp.event.record()
    grp.members = []
    o = 0
    for w, (all_t, pf_e, pf_r) in zip(windows, prepared):
        p = _Prepared()
        p.glist = glist_all[o:o + len(w)]
        o += len(w)
        p.new_graphs = []
        p.all_t, p.pf_ent, p.pf_rel = all_t, pf_e, pf_r
        p.sizes_host, p.event, p.totals = None, grp.event, None
        p.group = grp
        grp.members.append(p)
    return list(grp.members)


def _finish_prepare(p, filters=True):
    """Read the sizes of a prepared timestamp and finalise its graphs; with filters=False the two filter lists are left
    for _finish_filters (so that their fill launches can be enqueued BEHIND the evolution that only needs the graphs)."""
    from .graph import finish_sub_graphs
    if getattr(p, "group", None) is not None:
        p.group.finish()
        return _finish_filters(p) if filters else None
    p.event.synchronize()
    sizes = p.sizes_host.tolist()
    if p.new_graphs:
        finish_sub_graphs(p.new_graphs, [sizes[2 + 8 * i:10 + 8 * i] for i in range(len(p.new_graphs))])
    p.totals = (sizes[0], sizes[1])
    return _finish_filters(p) if filters else None


def _finish_filters(p):
    return utils.filter_lists_finish2(p.pf_ent, p.totals[0], p.pf_rel, p.totals[1])


@torch.no_grad()
def _test_multi_step(model, input_list, test_list, num_rels, num_nodes, static_graph, dev, topk, relation_evaluation,
                     return_ranks):
    """src/main.py:66-100 with args.multi_step: predict, rank, then replace the oldest history snapshot by the predicted
    one (top-k entities -- or relations with --relation-evaluation -- of every query)."""
    from .graph import build_sub_graph
    ranks = [[], [], [], []]
    gpu = dev.index if dev.index is not None else 0
    for snap in test_list:
        glist = [build_sub_graph(num_nodes, num_rels, g, True, gpu) for g in input_list]
        triples = torch.as_tensor(snap).to(dev)
        all_t, score, score_rel = model.predict(glist, num_rels, static_graph, triples, True)
        f_rel = utils.filter_csr_from_snapshot(all_t, num_nodes, 1)
        f_ent = utils.filter_csr_from_snapshot(all_t, 2 * num_rels, 0)
        _, _, rank_r, frank_r = utils.get_total_rank(all_t, score_rel, None, 1000, rel_predict=1, filter_csr=f_rel)
        _, _, rank, frank = utils.get_total_rank(all_t, score, None, 1000, rel_predict=0, filter_csr=f_ent)
        for lst, v in zip(ranks, (rank, frank, rank_r, frank_r)):
            lst.append(v.cpu())
        if relation_evaluation:
            predicted = utils.construct_snap_r(all_t, num_nodes, num_rels, score_rel, topk)
        else:
            predicted = utils.construct_snap(all_t, num_nodes, num_rels, score, topk)
        if len(predicted):
            input_list.pop(0)
            input_list.append(predicted)

    def mrr(lst):
        return float(torch.mean(1.0 / torch.cat(lst).float())) if lst else float("nan")
    out = (mrr(ranks[0]), mrr(ranks[1]), mrr(ranks[2]), mrr(ranks[3]))
    return (out, ranks) if return_ranks else out


def _check_time_aware_filters(test_list, num_rels, all_ans_list, all_ans_r_list):
    """test() filters with the test snapshot's own answers (what the reference's main passes: rgcn/utils.py:286-304).
    Filter dicts handed in must describe exactly those sets -- anything else (static / global filtering) would silently
    be ignored, so it is refused here; utils.get_total_rank(..., all_ans) takes arbitrary dicts."""
    for name, lst, rel_p in (("all_ans_list", all_ans_list, False), ("all_ans_r_list", all_ans_r_list, True)):
        if lst is None:
            continue
        if len(lst) != len(test_list):
            raise ValueError(f"regcn_b200.test: {name} has {len(lst)} entries for {len(test_list)} test snapshots")
        for k, (snap, given) in enumerate(zip(test_list, lst)):
            arr = snap.cpu().numpy() if torch.is_tensor(snap) else snap
            own = utils.load_all_answers_for_filter(arr, num_rels, rel_p)
            same = own.keys() == given.keys() and all(
                own[a].keys() == given[a].keys() and all(set(own[a][b]) == set(given[a][b]) for b in own[a]) for a in own)
            if not same:
                raise ValueError(f"regcn_b200.test: {name}[{k}] is not the time-aware answer set of test snapshot {k}; "
                                 "test() builds its filter lists from the snapshot itself (the reference's main does the "
                                 "same) -- rank with utils.get_total_rank(test_triples, score, all_ans, ...) for other filters")


def test(model, history_list, test_list, num_rels, num_nodes, use_cuda=True, all_ans_list=None, all_ans_r_list=None,
         model_name=None, static_graph=None, mode="eval", test_history_len=None, multi_step=False, device=None,
         return_ranks=False, topk=10, relation_evaluation=False):
    """Drop-in for the reference's evaluation loop `test()` (src/main.py:33-123, hyperbolic_main.py:60-170): slide a
    window of `test_history_len` snapshots over `test_list`, predict every entity / relation for each test snapshot,
    rank raw + time-filtered, return (mrr_raw, mrr_filter, mrr_raw_r, mrr_filter_r).

    What differs from the reference is only where the time goes:
      * a snapshot's edge index is built once (SnapshotCache) instead of L times;
      * the filter sets are the test snapshot's own answers (time-aware filtering, rgcn/utils.py:286-304), built on the
        device from the query triples -- `all_ans_list` / `all_ans_r_list` (the reference's nested dicts) are accepted
        for signature compatibility and, when given, must describe exactly those sets;
      * entity ranks come out of the scoring GEMM's counting epilogue (no (B,N) score matrix, no sort);
      * the loop is software-pipelined: while the GPU works on timestamp k, the host prepares timestamp k+1 (H2D copy,
        index build of the snapshot that just entered the window, filter-list counting) and only then waits for the few
        integers it needs to size step k+1's buffers; ranks travel back through pinned buffers, one copy per timestamp.
    `mode="test"` with `model_name` loads the checkpoint first like the reference.  `multi_step=True` feeds the model's
    own top-`topk` predictions back as the next history snapshot (src/main.py:90-97, --multi-step / --topk /
    --relation-evaluation); that recurrence is sequential, so it runs the plain (un-pipelined) loop on dense scores."""
    import os
    import time
    if not use_cuda:
        raise RuntimeError("regcn_b200.test: use_cuda=False is not supported (no CPU path)")
    from . import _lib
    from .graph import SnapshotCache
    _lib.require_device()
    if mode == "test" and model_name is not None:
        ck = torch.load(model_name, map_location="cpu")
        model.load_state_dict(ck["state_dict"] if "state_dict" in ck else ck)
    model.eval()
    _check_time_aware_filters(test_list, num_rels, all_ans_list, all_ans_r_list)
    dev = device if device is not None else next(model.parameters()).device
    L = test_history_len if test_history_len is not None else getattr(model, "sequence_len", len(history_list))
    input_list = [snap for snap in history_list[-L:]]
    if multi_step:
        return _test_multi_step(model, input_list, test_list, num_rels, num_nodes, static_graph, dev, topk,
                                relation_evaluation, return_ranks)
    # consecutive test timestamps are independent (ground-truth history, src/main.py:98-100): G of them share one
    # batched recurrence (timestamps_per_batch / RecurrentRGCN.forward_batch); their decode + rank stays per timestamp
    G = timestamps_per_batch(model, num_nodes, static_graph) if len(input_list) else 1
    PREP_DEPTH = 2 if G == 1 else 2 * G   # timestamps whose device-side preparation is in flight ahead of the one(s) evaluated
    AHEAD = 8           # test snapshots whose edge index is built per batched launch (one CTA each, concurrently): the
                        # one-CTA build of a single snapshot is 86 us of latency at the ICEWS18 size, eight cost the same
    cache = SnapshotCache(num_nodes, num_rels, dev, capacity=max(2 * L + 4, 10) + AHEAD + PREP_DEPTH + G)
    fused_ok = ops.gemm_impl() in ("tc", "tc1")
    # ConvTransE / ConvTransR in fp32-parity mode: decode + rank of a timestamp is one C call (same kernels, same ranks)
    one_call = (ops.gemm_impl() == "tc" and ops.score_dtype() == "fp32" and os.environ.get("REGCN_DECODE_ENGINE", "1") != "0"
                and type(model.decoder_ob).__name__ == "ConvTransE" and type(model.rdecoder).__name__ == "ConvTransR"
                and model.h_dim % 4 == 0 and model.h_dim <= 256)
    K = len(test_list)
    # pinned staging, allocated once: size slots (ring over the timestamps being run and the ones being prepared) and one
    # result area holding [rank | frank | rank_r | frank_r] of every timestamp
    n_slots = PREP_DEPTH + G + 1
    size_slots = torch.empty((n_slots, 2 + 8 * (L + AHEAD + 2)), dtype=torch.int32, pin_memory=True)
    offs = [0]
    for snap in test_list:
        offs.append(offs[-1] + 8 * int(snap.shape[0]))
    result_host = torch.empty(max(offs[-1], 1), dtype=torch.int32, pin_memory=True)
    results = []
    # The preparation of a timestamp (H2D copy, index build of the snapshot entering the window, filter-list counting, D2H
    # of the sizes) is enqueued on the SAME stream, PREP_DEPTH timestamps ahead: when the host is about to enqueue the next
    # group the sizes it needs were produced in front of the current group's kernels and have long arrived, so the host
    # never blocks on the GPU and the GPU never waits for the host (with a depth of one the GPU idled ~0.2 ms per step
    # while the host fetched the sizes and enqueued the next evolution).  A separate preparation stream would also overlap
    # those small kernels with the evolution, but makes the caching allocator fall back to fresh allocations for every
    # cross-stream tensor (measured: 2.7-4.7 ms per step); REGCN_PREP_STREAM=1 selects it anyway.
    # Non multi-step evaluation feeds ground-truth history (src/main.py:98-100), so every window is known up front.
    main_stream = torch.cuda.current_stream()
    prep_stream = main_stream
    if os.environ.get("REGCN_PREP_STREAM") == "1":
        prep_stream = torch.cuda.Stream(device=dev)
        prep_stream.wait_stream(main_stream)
    base = list(input_list)

    def window(k):
        return (base + list(test_list[:k]))[-len(base):] if len(base) else []

    def ahead(j):
        # test snapshot i joins the window at step i + 1; every AHEAD-th step builds the next AHEAD of them at once
        return list(test_list[j:min(j + AHEAD, K - 1)]) if (j % AHEAD == 0 and len(base)) else []

    def prepare(j):
        return _prepare(cache, window(j), test_list[j], num_rels, dev, size_slots[j % n_slots], prep_stream, main_stream,
                        ahead(j))

    # start-up: only the first group is prepared in front of the first evolution (every prepare costs ~0.2 ms of host time
    # during which the GPU has nothing to do); the following groups are prepared behind it, see the two top-ups below
    # Group sizes of the call: a short first group (8), then equal shares of the rest in groups of at most G.  The
    # preparation of a group is ~0.1 ms of host time per timestamp with nothing but the previous group's decodes for the
    # GPU to run meanwhile, and the evolution is cheaper per timestamp the more windows share a recurrence (ICEWS18 shape:
    # 259 / 209 / 184 us at 10 / 20 / 30 windows).  Measured, ms per timestamp of a 32-timestamp call (profiles/e2e_loop.py;
    # REGCN_TEST_RAMP = first group sizes): "4,8" (the ramp used before: groups 4, 8, 20) 0.774, "4,12" 0.767, "2,6" 0.761,
    # "4" 0.747, "6" 0.747, "8" (groups 8, 24) 0.743.
    sizes = group_sizes(K, G, os.environ.get("REGCN_TEST_RAMP", "8"))
    starts = [0]
    for n_ in sizes:
        starts.append(starts[-1] + n_)
    G_FIRST = sizes[0] if sizes else 1
    # A whole group is prepared by ONE batched call (_prepare_group; its sizes travel in one pinned slot of the ring):
    # 0.70-0.72 vs 0.743-0.756 ms per timestamp at the ICEWS18 shape with groups of 8 + 24 (30 calls, no outlier; it needs
    # the device allocations of the batch in size classes that repeat from call to call -- a fresh cudaMalloc while kernels
    # are queued stalls the host for tens of ms: DESIGN 11.2).  REGCN_PREP_BATCH=0: one preparation per timestamp.
    batched = G > 1 and prep_stream is main_stream and os.environ.get("REGCN_PREP_BATCH", "1") != "0"
    group_slots = torch.empty((4, 2 * 32 + 8 * (L + 32 + 2)), dtype=torch.int32, pin_memory=True) if batched else None
    n_ranges = [0]
    on_device = {}
    stage_all = torch.empty(3 * (offs[-1] // 8), dtype=torch.int64, pin_memory=True) if batched else None

    def prepare_range(a, b):
        members = None
        if batched:
            members = _prepare_group(cache, [window(j) for j in range(a, b)], list(test_list[a:b]), num_rels, dev,
                                     group_slots[n_ranges[0] % 4], stage_all[3 * (offs[a] // 8):3 * (offs[b] // 8)], on_device)
            n_ranges[0] += 1
        return members if members is not None else [prepare(j) for j in range(a, b)]

    queue = prepare_range(0, min(G_FIRST, K)) if G > 1 else [prepare(j) for j in range(min(PREP_DEPTH, K))]
    next_j = len(queue)
    _tm = os.environ.get("REGCN_TEST_TIMING") == "1"
    _acc = [0.0] * 6
    k = 0
    gi_ = 0
    while k < K:
        n_g = sizes[gi_]
        gi_ += 1
        group = [queue.pop(0) for _ in range(n_g)]
        _t0 = time.perf_counter()
        for cur in group:
            cur.event.synchronize()
            if prep_stream is not main_stream:
                main_stream.wait_event(cur.event)
        _t1 = time.perf_counter()
        for cur in group:
            _finish_prepare(cur, filters=False)
        _t2 = time.perf_counter()
        if n_g > 1:
            states = model.forward_batch([cur.glist for cur in group])
        else:
            evolve_embs, _, r_emb, _, _ = model.forward(group[0].glist, static_graph, True)
            states = [(evolve_embs[-1], r_emb)]
        filters = [_finish_filters(cur) for cur in group]      # the fill launches queue up behind the evolution
        _t3 = time.perf_counter()
        # the next group's preparation goes behind this evolution, interleaved with this group's decodes: every prepare
        # costs ~0.1 ms of host time, and with a decode (~0.4 ms of kernels) enqueued in front of each few of them the GPU
        # has work while the host issues them (at the start of a call nothing else is queued: 16 preparations in a row left
        # the GPU idle for ~2 ms).  Their sizes are on the host when the next iteration asks for them (in steady state they
        # were already prepared one iteration earlier).
        prep_target = min(K, starts[min(gi_ + 1, len(sizes))]) if G > 1 else min(K, k + n_g)   # through the next group
        prep_each = -(-max(0, prep_target - next_j) // max(1, n_g // 2))       # all of them within the first half of the decodes:
        # the host asks for their sizes at the top of the next iteration, when the GPU still has the second half queued
        _t3b = time.perf_counter()
        _tprep = 0.0
        for i, (cur, (emb, r_emb), (f_ent, f_rel)) in enumerate(zip(group, states, filters)):
            all_t = cur.all_t
            if one_call:
                packed = convtrans_decode_rank(model, emb, r_emb, all_t, f_ent, f_rel)
            else:
                if model.layer_norm:
                    if hasattr(model, "_c_float"):
                        emb = ops.row_map(emb, ops.ROW_TANGENT_NORMALIZE, c=model._c_float)
                    else:
                        emb = ops.row_map(emb, ops.ROW_NORMALIZE)
                if fused_ok:
                    q, cand, hyp, col_bias = _scoring_operands(model, emb, r_emb, all_t)
                    target = all_t[:, 2].to(torch.int32).contiguous()
                    pa, pe = f_ent.pairs(target)
                    raw, filt, _ = ops.fused_rank_counts(q, cand, target, f_ent.ptr, f_ent.idx, pa, pe, hyp=hyp,
                                                         col_bias=col_bias, filt_end=f_ent.end)
                    rank, frank = ops.counts_to_ranks(raw, filt)
                else:
                    score = model.decoder_ob.forward(emb, r_emb, all_t, mode="test")
                    _, _, rank, frank = utils.get_total_rank(all_t, score, None, 1000, rel_predict=0, filter_csr=f_ent)
                score_rel = model.rdecoder.forward(emb, r_emb, all_t, mode="test")
                raw_r, filt_r, _ = ops.rank_dense(score_rel, all_t, 1, f_rel.ptr, f_rel.idx, filt_end=f_rel.end)
                rank_r, frank_r = ops.counts_to_ranks(raw_r, filt_r)
                packed = torch.cat((rank, frank, rank_r, frank_r)).to(torch.int32)
            host = result_host[offs[k + i]:offs[k + i + 1]]
            host.copy_(packed, non_blocking=True)                      # this timestamp's result, device -> host
            results.append(host)
            _tp0 = time.perf_counter()
            if batched:
                # the next group, as one batch, behind this group's first decode (the GPU has the evolution's tail and a
                # decode queued while the host issues it; its sizes are on the host long before the next iteration)
                if next_j < prep_target:
                    queue.extend(prepare_range(next_j, prep_target))
                    next_j = prep_target
            else:
                for _ in range(prep_each):
                    if next_j < prep_target:
                        queue.append(prepare(next_j))
                        next_j += 1
            _tprep += time.perf_counter() - _tp0
        while next_j < prep_target:
            queue.append(prepare(next_j))
            next_j += 1
        _t4 = time.perf_counter()
        k += n_g
        # the window slides (src/main.py:98-100); prepare the timestamps up to PREP_DEPTH ahead while the GPU is busy
        while G == 1 and next_j < min(K, k + PREP_DEPTH):
            queue.append(prepare(next_j))
            next_j += 1
        _t5 = time.perf_counter()
        for _i, _d in enumerate((_t1 - _t0, _t2 - _t1, _t3 - _t2, _t4 - _t3b - _tprep, _t5 - _t4 + _t3b - _t3 + _tprep)):
            _acc[_i] += _d
    if _tm and K:
        print("test() host ms/step: wait %.3f finish %.3f forward %.3f decode+rank %.3f prepare %.3f" %
              tuple(1e3 * a / K for a in _acc[:5]))
    torch.cuda.current_stream().synchronize()
    ranks = [[], [], [], []]
    for host in results:
        B = host.numel() // 4
        for j in range(4):
            ranks[j].append(host[j * B:(j + 1) * B].long())
    def mrr(lst):
        if not lst:
            return float("nan")
        return float(torch.mean(1.0 / torch.cat(lst).float()))
    out = (mrr(ranks[0]), mrr(ranks[1]), mrr(ranks[2]), mrr(ranks[3]))
    if return_ranks:
        return out, ranks
    return out
