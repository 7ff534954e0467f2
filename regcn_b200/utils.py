"""Evaluation utilities with the reference's names and return values (rgcn/utils.py).

`get_total_rank` (rgcn/utils.py:136-166) keeps its signature and returns; the two full sorts and the
per-query python filter loop (utils.py:21-25, 51-75) are replaced by the count-based rank kernels
(regcn_b200/csrc/rank.cu).  Host-side helpers that only shape python data (answer dictionaries,
`split_by_time`) are restated here so callers of the reference find them.
"""
import numpy as np
import torch

from . import ops
from .graph import build_sub_graph  # noqa: F401  (same import site as the reference: utils.build_sub_graph)


# ------------------------------------------------------------------------------ answer sets
def _add(d, k1, k2, v):
    d.setdefault(k1, {}).setdefault(k2, set()).add(v)


def load_all_answers_for_filter(total_data, num_rel, rel_p=False):
    """rgcn/utils.py:264-283: {e1: {r: {e2}}} incl. inverse queries, or {e1: {e2: {r}}} when rel_p."""
    all_ans = {}
    for line in total_data:
        s, r, o = (int(x) for x in line[:3])
        if rel_p:
            _add(all_ans, s, o, r)
            _add(all_ans, o, s, r + num_rel)
        else:
            _add(all_ans, o, r + num_rel, s)
            _add(all_ans, s, r, o)
    return all_ans


def split_by_time(data):
    """rgcn/utils.py:306-339 (without the sanity print): list of (T_i,3) arrays, one per timestamp."""
    data = np.asarray(data)
    snapshot_list, snapshot, latest_t = [], [], 0
    for i in range(len(data)):
        t = data[i][3]
        if latest_t != t:
            latest_t = t
            if len(snapshot):
                snapshot_list.append(np.array(snapshot).copy())
            snapshot = []
        snapshot.append(data[i][:3])
    if len(snapshot) > 0:
        snapshot_list.append(np.array(snapshot).copy())
    return snapshot_list


def load_data(dataset, data_dir="../data"):
    """rgcn/utils.py:356-365 (temporal datasets)."""
    from . import knowledge_graph as knwlgrh
    return knwlgrh.load_data(dataset, data_dir)


def load_all_answers_for_time_filter(total_data, num_rels, num_nodes, rel_p=False):
    """rgcn/utils.py:286-304."""
    return [load_all_answers_for_filter(snap, num_rels, rel_p) for snap in split_by_time(total_data)]


# ------------------------------------------------------------------------------ filter CSR
class FilterCSR:
    """Per-query sorted ids of the true answers (the reference's all_ans[h][r] / all_ans[h][t] sets) on the device:
    the list of query b is idx[ptr[b] : end[b]] (end=None means the compact CSR form idx[ptr[b] : ptr[b+1]])."""

    def __init__(self, ptr_t, idx_t, end_t=None, pairs=None):
        self.ptr, self.idx, self.end = ptr_t, idx_t, end_t
        self._pairs = pairs

    def pairs(self, target):
        """(pair_a, pair_e) int32 lists for the fused rank path: the B (query, target) pairs followed by one
        (query, candidate) pair per idx slot."""
        if self._pairs is None:
            B = self.ptr.numel() - 1
            counts = (self.ptr[1:] - self.ptr[:-1]).long()
            rows = torch.repeat_interleave(torch.arange(B, device=self.ptr.device), counts)
            n = int(rows.numel())
            pa = torch.cat((torch.arange(B, device=self.ptr.device), rows)).to(torch.int32)
            pe = torch.cat((target.to(torch.int32), self.idx[:n])).contiguous()
            self._pairs = (pa.contiguous(), pe)
        return self._pairs

    def lists(self):
        """Host copy as python lists (tests / debugging)."""
        p, i = self.ptr.cpu().tolist(), self.idx.cpu().tolist()
        e = self.end.cpu().tolist() if self.end is not None else p[1:]
        return [i[p[b]:e[b]] for b in range(len(e))]


class _PendingFilter:
    """Phase 1 of the kernel-built filter lists: match counts and list offsets are on the device, the slot total has
    not been read back yet (so that several index builds can share one host synchronisation)."""

    def __init__(self, all_triples, rel_predict):
        from ._lib import call, ptr
        self.triples = all_triples
        self.B = B = all_triples.shape[0]
        dev = all_triples.device
        self.key_col, self.ans_col = (2, 1) if rel_predict else (1, 2)
        counts = torch.empty(B, device=dev, dtype=torch.int32)
        call("regcn_filter_count", ptr(all_triples), B, self.key_col, ptr(counts))
        csum = torch.cumsum(counts, 0, dtype=torch.int32)
        self.beg = csum - counts
        self.total = csum[-1:] if B else torch.zeros(1, device=dev, dtype=torch.int32)

    def finish(self, total=None):
        from ._lib import call, ptr
        B, dev = self.B, self.triples.device
        total = int(self.total.item()) if total is None else int(total)
        idx = torch.empty(max(total, 1), device=dev, dtype=torch.int32)
        end = torch.empty(B, device=dev, dtype=torch.int32)
        pa = torch.empty(B + total, device=dev, dtype=torch.int32)
        pe = torch.empty(B + total, device=dev, dtype=torch.int32)
        call("regcn_filter_fill", ptr(self.triples), B, self.key_col, self.ans_col, ptr(self.beg), ptr(idx), ptr(end),
             ptr(pa), ptr(pe))
        return FilterCSR(self.beg, idx, end, pairs=(pa, pe))


def filter_lists_finish2(pf_ent, total_ent, pf_rel, total_rel):
    """finish() of the entity and the relation filter of the same queries as ONE launch (`regcn_filter_fill2`); all
    outputs live in one allocation.  Returns (FilterCSR entity, FilterCSR relation), identical to the two finish() calls."""
    from ._lib import call, ptr
    if pf_ent.triples is not pf_rel.triples or pf_ent.key_col != 1 or pf_rel.key_col != 2:
        return pf_ent.finish(total_ent), pf_rel.finish(total_rel)
    B, dev = pf_ent.B, pf_ent.triples.device
    te, tr = int(total_ent), int(total_rel)
    ne, nr = max(te, 1), max(tr, 1)
    buf = torch.empty(ne + nr + 2 * B + 2 * (B + te) + 2 * (B + tr), device=dev, dtype=torch.int32)
    parts, o = [], 0
    for n in (ne, B, B + te, B + te, nr, B, B + tr, B + tr):
        parts.append(buf[o:o + n])
        o += n
    idx_e, end_e, pa_e, pe_e, idx_r, end_r, pa_r, pe_r = parts
    call("regcn_filter_fill2", ptr(pf_ent.triples), B, ptr(pf_ent.beg), ptr(idx_e), ptr(end_e), ptr(pa_e), ptr(pe_e),
         ptr(pf_rel.beg), ptr(idx_r), ptr(end_r), ptr(pa_r), ptr(pe_r))
    return (FilterCSR(pf_ent.beg, idx_e, end_e, pairs=(pa_e, pe_e)), FilterCSR(pf_rel.beg, idx_r, end_r, pairs=(pa_r, pe_r)))


def filter_lists_begin(all_triples, rel_predict=0):
    return _PendingFilter(all_triples.contiguous(), rel_predict)


def queries_prepare(test_triples, num_rels):
    """Device-side preparation of one test snapshot in ONE C call (`regcn_queries_prepare`): returns (all_triples,
    pending_entity_filter, pending_relation_filter, totals) -- all_triples = the (T,3) int64 device triples followed by
    their inverses (src/rrgcn.py:184-186), the two filter objects are in the state filter_lists_begin leaves them in
    (counts and offsets on the device), totals (2,) int32 = the slot totals their finish() needs from the host."""
    from ._lib import call, ptr
    t = test_triples.contiguous()
    T = int(t.shape[0])
    B = 2 * T
    dev = t.device
    all_t = torch.empty((B, 3), device=dev, dtype=torch.int64)
    work = torch.empty(4 * B + 2, device=dev, dtype=torch.int32)          # counts (2,B) | offsets (2,B) | totals (2)
    counts, beg, totals = work[:2 * B].view(2, B), work[2 * B:4 * B].view(2, B), work[4 * B:]
    call("regcn_queries_prepare", ptr(t), T, int(num_rels), ptr(all_t), ptr(counts), ptr(beg), ptr(totals))
    out = []
    for rel_predict in (0, 1):
        pf = _PendingFilter.__new__(_PendingFilter)
        pf.triples, pf.B = all_t, B
        pf.key_col, pf.ans_col = (2, 1) if rel_predict else (1, 2)
        pf.beg = beg[rel_predict]
        pf.total = totals[rel_predict:rel_predict + 1]
        out.append(pf)
    return all_t, out[0], out[1], totals


def queries_prepare_batch(triples_cat, Ts, num_rels):
    """`queries_prepare` for the test snapshots of a GROUP of timestamps in one C call / three launches
    (`regcn_queries_prepare_batch`).  triples_cat: (sum Ts, 3) int64 device tensor, the snapshots back to back; Ts: their
    row counts (all > 0, at most 32 of them).  Returns ([(all_triples, pending_entity_filter, pending_relation_filter)],
    totals (n, 2) int32 on the device) -- member g is what queries_prepare returns for snapshot g alone."""
    import ctypes
    from ._lib import call, ptr
    n = len(Ts)
    toff = [0]
    for T in Ts:
        toff.append(toff[-1] + int(T))
    tot = toff[-1]
    dev = triples_cat.device
    t = triples_cat.contiguous()
    cap = (tot + 65535) // 65536 * 65536         # allocation sizes that repeat from call to call (caching allocator)
    all_t = torch.empty((2 * cap, 3), device=dev, dtype=torch.int64)[:2 * tot]
    work = torch.empty(8 * cap + 64, device=dev, dtype=torch.int32)      # counts (4 tot) | offsets (4 tot) | totals (n, 2)
    counts, beg, totals = work[:4 * tot], work[4 * tot:8 * tot], work[8 * tot:8 * tot + 2 * n].view(n, 2)
    toff_c = (ctypes.c_int32 * (n + 1))(*toff)
    call("regcn_queries_prepare_batch", ptr(t), ctypes.cast(toff_c, ctypes.c_void_p), n, int(num_rels), ptr(all_t), ptr(counts),
         ptr(beg), ptr(totals))
    out = []
    for g in range(n):
        B = 2 * int(Ts[g])
        a = all_t[2 * toff[g]:2 * toff[g + 1]]
        bg = beg[4 * toff[g]:4 * toff[g + 1]].view(2, B)
        pfs = []
        for rel_predict in (0, 1):
            pf = _PendingFilter.__new__(_PendingFilter)
            pf.triples, pf.B = a, B
            pf.key_col, pf.ans_col = (2, 1) if rel_predict else (1, 2)
            pf.beg = bg[rel_predict]
            pf.total = totals[g, rel_predict:rel_predict + 1]
            pfs.append(pf)
        out.append((a, pfs[0], pfs[1]))
    return out, totals


def filter_lists_from_queries(all_triples, rel_predict=0):
    """Kernel-built time-aware filter lists for the queries themselves (the test snapshot incl. inverses): two
    launches + one scan + one host read of the slot total; also yields the fused-rank pair lists.  All-pairs scan
    (one warp per query), meant for the few thousand queries of a timestamp."""
    return _PendingFilter(all_triples.contiguous(), rel_predict).finish()


def filter_csr_from_dict(test_triples, all_ans, rel_predict=0, device=None):
    """Build the CSR from the reference's nested dict (one pass over the queries on the host)."""
    tt = test_triples.detach().cpu().numpy()
    ptr_l, idx_l = [0], []
    for h, r, t in tt:
        key2 = int(t) if rel_predict else int(r)
        ans = sorted(all_ans[int(h)][key2])
        idx_l.extend(ans)
        ptr_l.append(len(idx_l))
    dev = device if device is not None else test_triples.device
    return FilterCSR(torch.tensor(ptr_l, dtype=torch.int32, device=dev),
                     torch.tensor(idx_l if idx_l else [0], dtype=torch.int32, device=dev))


def filter_csr_from_snapshot(all_triples, num_keys2, rel_predict=0, num_answers=None, use_kernels=None):
    """Vectorised, device-side equivalent of load_all_answers_for_filter + per-query lookup when the filter set is
    'every answer among these queries themselves' (time-aware filtering, rgcn/utils.py:286-304: the test snapshot's own
    triples incl. inverses = exactly `all_triples` of predict()).  key = (h, r) -> answers t (entity prediction) or
    (h, t) -> answers r (relation prediction)."""
    if use_kernels is None:
        use_kernels = all_triples.is_cuda and 0 < all_triples.shape[0] <= 32768
    if use_kernels:
        return filter_lists_from_queries(all_triples.contiguous(), rel_predict)
    h = all_triples[:, 0]
    k2 = all_triples[:, 2] if rel_predict else all_triples[:, 1]
    a = all_triples[:, 1] if rel_predict else all_triples[:, 2]
    key = h * int(num_keys2) + k2
    # unique (key, answer) pairs sorted by key then answer
    # `num_answers` (an upper bound on the answer ids: num_ents, or 2*num_rels for relation prediction) avoids a sync
    big = int(num_answers) if num_answers is not None else (int(a.max().item()) + 1 if a.numel() else 1)
    pair = torch.unique(key * big + a)
    pkey, pans = pair // big, pair % big
    ukey, counts = torch.unique_consecutive(pkey, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.searchsorted(ukey, key)
    q_start, q_cnt = starts[pos], counts[pos]
    ptr_t = torch.zeros(key.numel() + 1, dtype=torch.int64, device=key.device)
    ptr_t[1:] = torch.cumsum(q_cnt, 0)
    total = int(ptr_t[-1].item())
    rep = torch.repeat_interleave(torch.arange(key.numel(), device=key.device), q_cnt)
    within = torch.arange(total, device=key.device) - ptr_t[:-1][rep]
    idx_t = pans[q_start[rep] + within]
    return FilterCSR(ptr_t.to(torch.int32), idx_t.to(torch.int32) if total else torch.zeros(1, dtype=torch.int32, device=key.device))


# ------------------------------------------------------------------------------ ranking
def sort_and_rank(score, target):
    """rgcn/utils.py:21-25: 0-based rank of `target` in each row (count-based, stable-sort tie rule)."""
    trip = torch.zeros((score.shape[0], 3), dtype=torch.int64, device=score.device)
    trip[:, 2] = target
    raw, _, _ = ops.rank_dense(score.contiguous() if score.stride(1) != 1 else score, trip, 2)
    return raw.long()


def get_total_rank(test_triples, score, all_ans, eval_bz, rel_predict=0, filter_csr=None):
    """rgcn/utils.py:136-166.  Returns (filter_mrr, mrr, rank, filter_rank) with 1-based int64 ranks.

    `all_ans` is the reference's nested dict (or None); pass `filter_csr` to skip the host-side dict walk.
    Like the reference, the filtered entries of `score` are overwritten with -10000000 in place (utils.py:60,74)."""
    target_col = {0: 2, 1: 1, 2: 0}[rel_predict]
    test_triples = test_triples.contiguous()
    if filter_csr is None and all_ans is not None:
        filter_csr = filter_csr_from_dict(test_triples, all_ans, rel_predict=1 if rel_predict else 0,
                                          device=score.device)
    fp = filter_csr.ptr if filter_csr is not None else None
    fi = filter_csr.idx if filter_csr is not None else None
    fe = filter_csr.end if filter_csr is not None else None
    raw, filt, _ = ops.rank_dense(score, test_triples, target_col, fp, fi, filt_end=fe)
    rank, filter_rank = ops.counts_to_ranks(raw, filt)
    if filter_csr is not None:
        ops.apply_filter_(score, test_triples, target_col, fp, fi, filt_end=fe)
    mrr = torch.mean(1.0 / rank.float())
    filter_mrr = torch.mean(1.0 / filter_rank.float())
    return filter_mrr.item(), mrr.item(), rank, filter_rank


def stat_ranks(rank_list, method, verbose=True):
    """rgcn/utils.py:169-178."""
    total_rank = torch.cat(rank_list)
    mrr = torch.mean(1.0 / total_rank.float())
    if verbose:
        print("MRR ({}): {:.6f}".format(method, mrr.item()))
        for hit in (1, 3, 10):
            print("Hits ({}) @ {}: {:.6f}".format(method, hit, torch.mean((total_rank <= hit).float()).item()))
    return mrr


# ------------------------------------------------------------------------------ multi-step inference
def _construct(test_triples, num_rels, final_score, topK, rel_mode):
    from . import _lib
    from ._lib import call, ptr
    _lib.require_device()
    if not final_score.is_cuda:
        raise RuntimeError("regcn_b200: kernels take CUDA tensors only (no CPU fallback)")
    B, N = final_score.shape
    K = int(topK)
    test_triples = test_triples.to(final_score.device).contiguous()
    if final_score.stride(1) != 1:
        final_score = final_score.contiguous()
    top = torch.empty((B, K), device=final_score.device, dtype=torch.int32)
    out = torch.empty((B * K, 3), device=final_score.device, dtype=torch.int64)
    call("regcn_topk_construct_snap", final_score.data_ptr(), final_score.stride(0), B, N, K, ptr(test_triples),
         int(num_rels), rel_mode, ptr(top), ptr(out))
    return out


def construct_snap(test_triples, num_nodes, num_rels, final_score, topK):
    """rgcn/utils.py:367-381: the predicted snapshot from the top-K entities of every query (a device int64 (B*K,3)
    tensor where the reference returns a numpy array; `build_sub_graph` takes either).  Ties: ascending id."""
    return _construct(test_triples, num_rels, final_score, topK, 0)


def construct_snap_r(test_triples, num_nodes, num_rels, final_score, topK):
    """rgcn/utils.py:383-405: the same from the top-K relations."""
    return _construct(test_triples, num_rels, final_score, topK, 1)
