"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path (shard bounds, rank merge, timestamp slices)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from regcn_b200 import dist as rd
        rng = np.random.default_rng(0)
        B, N = 37, 101
        score = np.round(rng.standard_normal((B, N)).astype(np.float32) * 4) / 4     # ties on purpose
        target = rng.integers(0, N, B)
        filt = [sorted(set(rng.integers(0, N, 3).tolist()) | {int(target[b])}) for b in range(B)]
        lo, hi = rd.shard_bounds(N, rank, ws)
        # per-shard counts exactly as the rank kernel defines them (restated on the host)
        st = score[np.arange(B), target]
        raw = np.zeros(B, dtype=np.int32)
        fil = np.zeros(B, dtype=np.int32)
        for b in range(B):
            for j in range(lo, hi):
                if j == target[b]:
                    continue
                s = score[b, j]
                beats = s > st[b] or (s == st[b] and j < target[b])
                raw[b] += beats
                if j in filt[b]:
                    s = np.float32(restate.FILTER_SCORE)
                    beats = s > st[b] or (s == st[b] and j < target[b])
                fil[b] += beats
        rank_t, frank_t = rd.merge_counts(torch.from_numpy(raw), torch.from_numpy(fil))
        # single-process truth
        fscore = score.copy()
        for b in range(B):
            for j in filt[b]:
                if j != target[b]:
                    fscore[b, j] = restate.FILTER_SCORE
        ok = (np.array_equal(rank_t.numpy(), restate.stable_rank0(score, target) + 1)
              and np.array_equal(frank_t.numpy(), restate.stable_rank0(fscore, target) + 1))
        # timestamp data parallelism: slices tile the range, gather restores global order
        lo_t, hi_t = rd.timestamp_slice(7, rank, ws)
        local = torch.arange(lo_t, hi_t, dtype=torch.int64) * 10
        allr = rd.gather_ranks(local)
        ok = ok and torch.equal(allr, torch.arange(7, dtype=torch.int64) * 10)
        # query-sharded tower: row shards of a (B, d) matrix gathered back in order, ragged (37 = 19 + 18) and even sizes
        for rows in (37, 36, 1):
            full = torch.arange(rows * 5, dtype=torch.float32).view(rows, 5)
            b0, b1 = rd.shard_bounds(rows, rank, ws)
            ok = ok and torch.equal(rd.gather_rows(full[b0:b1].clone(), rows), full)
        # shard bounds follow the GROUP the collectives run on (a sub-group of one rank owns every candidate)
        g0, g1 = dist.new_group([0]), dist.new_group([1])
        mine = g0 if rank == 0 else g1
        ok = ok and rd.world(mine) == (0, 1) and rd.shard_bounds(N, *rd.world(mine)) == (0, N)
        q.put((rank, bool(ok), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_rank_merge_and_timestamp_dp_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == (0, 51) and res[1][2] == (51, 101)


def test_shard_bounds_cover_everything():
    from regcn_b200 import dist as rd
    for n in (0, 1, 7, 23033, 1_000_000):
        for ws in (1, 2, 3, 4, 8):
            spans = [rd.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
