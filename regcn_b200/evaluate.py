"""Evaluation step of the hot path as one call: evolve the history, score every entity, rank raw + filtered.

This is the loop body of the reference's test() (src/main.py:67-74, hyperbolic_main.py:100-113) for entity
prediction, expressed on device-resident inputs so that nothing but kernels runs inside it.
"""
import torch

from . import ops, utils


def _scoring_operands(model, emb, r_emb, all_triples):
    """(q, cand, hyp, col_bias) of the entity decoder: everything the scoring GEMM needs except the GEMM itself."""
    dec = model.decoder_ob
    name = type(dec).__name__
    if name == "ConvTransE":
        e_all, q = dec.query(emb, r_emb, all_triples)
        return q, e_all, None, None
    if name == "HyperbolicConvTransE":
        et = ops.row_map(emb, ops.ROW_LEAKY_TANH_LOG0, c=dec.c)
        q = dec._tower(et, r_emb.contiguous(), all_triples, 0, 1, always_bn2=False)
        return q, et, None, dec.b.detach()
    if name in ("HyperbolicRotH", "HyperbolicMuRP"):
        q, qss = dec.query(emb, r_emb, all_triples)
        cand = emb.contiguous()
        # entity_bias[subject] shifts a whole row and cannot change a rank: only the candidate bias is needed here
        return q, cand, dec.hyp_operands(qss, cand, all_triples), dec._entity_bias(all_triples)[0]
    raise NotImplementedError(name)


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
    mark("score", 0)
    emb = evolve_embs[-1]
    if model.layer_norm:
        if hasattr(model, "_c_float"):
            emb = ops.row_map(emb, ops.ROW_TANGENT_NORMALIZE, c=model._c_float)
        else:
            emb = ops.row_map(emb, ops.ROW_NORMALIZE)
    if fused:
        q, cand, hyp, col_bias = _scoring_operands(model, emb, r_emb, all_triples)
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
def score_rank_sharded(model, emb, r_emb, all_triples, filter_csr, group=None):
    """Entity-partitioned scoring + rank merge: this rank counts over its contiguous slice of the
    (replicated, already evolved) entity table with the fused kernel; ONE all_reduce(SUM) of the (2,B) int32 counts
    merges the shards.  Target and filter-entry scores come from the pair pass on the replicated table, so no score
    exchange is needed.  Returns (rank, filter_rank), identical on every rank."""
    from . import dist as rdist
    r, ws = rdist.world()
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
    inv = test[:, [2, 1, 0]]
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
