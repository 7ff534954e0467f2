"""Multi-GPU plumbing for the two places the path shards.  One process per GPU,
torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests).

1. Entity-sharded all-entity scoring + rank merge: candidate rows of the (activated) entity table are split
   contiguously over ranks; every rank scores all B queries against its shard and counts, per query, the
   candidates that beat the target (raw and time-filtered); ONE all_reduce(SUM) of a (2,B) int32 tensor merges the
   shards (rank = 1 + total count), after an all_reduce(SUM) of the (B,) target scores published by the owner shards.
   The query tower is sharded the other way round: every rank builds the queries of its B/G slice and one all_gather
   of the (B,d) query matrix (gather_rows) precedes the count -- at the ICEWS18 size the tower is half of a timestamp's
   decode time, so leaving it replicated caps the speed-up at ~1.
2. Query-timestamp data parallelism for evaluation: test timestamps are independent units (src/main.py:98-100,
   non multi-step); ranks take contiguous slices and one all_gather of the rank vectors closes the job.

Nothing here touches the kernels' arithmetic; the collectives move integer counts / ranks only.
"""
import torch
import torch.distributed as dist


def world(group=None):
    """(rank, world size) INSIDE `group` (None = the default group): shard bounds and the collectives that merge them
    must be taken from the same group, or the shards do not tile [0, N)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous [lo, hi) slice of n candidates for `rank`: sizes differ by at most one, earlier ranks larger."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def timestamp_slice(n_timestamps, rank, world_size):
    """Contiguous slice of test timestamps evaluated by `rank` (same rule as shard_bounds)."""
    return shard_bounds(n_timestamps, rank, world_size)


def merge_counts(raw_count, filt_count, group=None):
    """all_reduce(SUM) the per-shard 'beats the target' counts -> global 1-based ranks (int64)."""
    packed = torch.stack((raw_count.to(torch.int32), filt_count.to(torch.int32)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[0].long() + 1, packed[1].long() + 1


def gather_ranks(local_ranks, group=None):
    """all_gather variable-length rank vectors of the per-rank timestamp slices; returns the concatenation in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_ranks
    ws = dist.get_world_size(group)
    n = torch.tensor([local_ranks.numel()], device=local_ranks.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros(mx, device=local_ranks.device, dtype=local_ranks.dtype)
    pad[: local_ranks.numel()] = local_ranks
    bufs = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def gather_rows(local_rows, n_total, group=None):
    """all_gather the row blocks of a matrix that was computed in contiguous row shards (shard_bounds(n_total, rank, ws)):
    returns the full (n_total, d) matrix on every rank.  Shards are padded to the largest size for the collective (sizes
    differ by at most one row) and the padding rows are dropped afterwards."""
    r, ws = world(group)
    if ws == 1:
        return local_rows
    d = local_rows.shape[1]
    per = -(-int(n_total) // ws)
    pad = local_rows
    if local_rows.shape[0] != per:
        pad = torch.zeros((per, d), device=local_rows.device, dtype=local_rows.dtype)
        pad[: local_rows.shape[0]] = local_rows
    if local_rows.is_cuda:
        out = torch.empty((ws * per, d), device=local_rows.device, dtype=local_rows.dtype)
        dist.all_gather_into_tensor(out, pad.contiguous(), group=group)          # one NCCL all-gather, no staging copies
        if n_total % ws == 0:
            return out
        keep = torch.cat([torch.arange(i * per, i * per + shard_bounds(n_total, i, ws)[1] - shard_bounds(n_total, i, ws)[0],
                                       device=out.device) for i in range(ws)])
        return out.index_select(0, keep)
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    if n_total % ws == 0:
        return torch.cat(bufs)
    return torch.cat([b[: shard_bounds(n_total, i, ws)[1] - shard_bounds(n_total, i, ws)[0]] for i, b in enumerate(bufs)])


def sharded_score_rank(n_cand, triples, target_col, filter_csr, score_fn, group=None):
    """Entity-sharded scoring + rank merge.  `score_fn(lo, hi)` returns this rank's dense (B, hi-lo) score block
    for candidate rows [lo, hi) (every rank holds the replicated, evolved entity table).

    Exchange 1: the shard that owns a query's target publishes its score  -> all_reduce(SUM) of (B,) fp32
                (exactly the value sitting in the dense block, so counts compare like with like);
    Exchange 2: per-shard raw / filtered 'beats the target' counts        -> all_reduce(SUM) of (2,B) int32.
    Returns (rank, filter_rank), identical on all ranks."""
    from . import ops
    from ._lib import call, ptr
    r, ws = world(group)
    lo, hi = shard_bounds(n_cand, r, ws)
    block = score_fn(lo, hi)
    B = block.shape[0]
    tscore = torch.zeros(B, device=block.device, dtype=torch.float32)
    call("regcn_gather_target_score", block.data_ptr(), block.stride(0), B, hi - lo, ptr(triples), target_col, lo,
         ptr(tscore))
    if ws > 1:
        dist.all_reduce(tscore, op=dist.ReduceOp.SUM, group=group)
    fp = filter_csr.ptr if filter_csr is not None else None
    fi = filter_csr.idx if filter_csr is not None else None
    fe = filter_csr.end if filter_csr is not None else None
    raw, filt, _ = ops.rank_dense(block, triples, target_col, fp, fi, col_offset=lo, target_score=tscore, filt_end=fe)
    if filt is None:
        filt = raw
    return merge_counts(raw, filt, group)
