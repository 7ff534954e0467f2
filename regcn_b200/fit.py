"""The reference's training epoch (src/main.py:213-246, hyperbolic_main.py:547-628) as a library call.

    for train_sample_num in shuffled(range(len(train_list))):          # index 0 is skipped
        input_list = train_list[max(0, t - train_history_len) : t]
        history_glist = [build_sub_graph(...) for snap in input_list]   # rebuilt every step in the reference
        loss_e, loss_r, loss_static[, loss_radius] = model.get_loss(history_glist, train_list[t], static_graph, use_cuda)
        loss = task_weight*loss_e + (1-task_weight)*loss_r + loss_static [+ loss_radius]
        loss.backward(); clip_grad_norm_(grad_norm); optimizer.step(); optimizer.zero_grad()

Host-side orchestration only (every arithmetic step is the kernels behind `get_loss` / `optim.Adam`).  What differs from
the reference is where the time goes: a snapshot's edge index is built once and kept resident (`SnapshotCache`) instead of
`train_history_len` times per step, the snapshots travel host->device once, and the per-step losses stay on the device
until the end of the epoch (one synchronisation per epoch instead of four `.item()` per step).
`triple_batch_size` reproduces hyperbolic_main.py:585-598: the snapshot's triples are cut into mini-batches, `get_loss`
(and with it the whole evolution) runs per mini-batch, gradients accumulate, one optimiser step per snapshot.
"""
import random

import numpy as np
import torch

from . import optim as _optim
from .graph import SnapshotCache, finish_sub_graphs


def fit_epoch(model, optimizer, train_list, num_rels, num_nodes, train_history_len, task_weight=0.7, grad_norm=1.0,
              static_graph=None, shuffle=True, triple_batch_size=None, device=None, cache=None, order=None):
    """One epoch over `train_list` (list of (T_i,3) int arrays / tensors, one per timestamp).  Returns a dict with the
    mean total / entity / relation / static (/ radius) losses of the epoch (python floats) and the per-step totals."""
    dev = device if device is not None else next(model.parameters()).device
    model.train()
    cache = cache if cache is not None else SnapshotCache(num_nodes, num_rels, dev, capacity=max(len(train_list) + 2, 8))
    idx = list(range(len(train_list))) if order is None else list(order)
    if shuffle and order is None:
        random.shuffle(idx)
    # device copies of the output snapshots, made once
    dev_snaps = {}
    totals, parts = [], []
    for t in idx:
        if t == 0:
            continue                                                   # src/main.py:224
        input_list = train_list[max(0, t - train_history_len):t]
        glist, new = cache.ensure(input_list)
        if new:
            finish_sub_graphs(new)
        out = dev_snaps.get(t)
        if out is None:
            snap = train_list[t]
            out = (snap if isinstance(snap, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(snap, dtype=np.int64)))
            out = dev_snaps[t] = out.to(dev).long()
        batches = [out] if not triple_batch_size else [out[b:b + triple_batch_size] for b in
                                                       range(0, out.shape[0], triple_batch_size)]
        step_parts = None
        for mb in batches:
            if mb.shape[0] < 1:
                continue
            losses = model.get_loss(glist, mb, static_graph, True)
            loss = task_weight * losses[0] + (1 - task_weight) * losses[1]
            for extra in losses[2:]:
                loss = loss + extra
            loss.backward()
            det = torch.stack([l.detach().reshape(-1)[0] for l in losses])
            step_parts = det if step_parts is None else step_parts + det
        if step_parts is None:
            continue
        step_parts = step_parts / len(batches)
        _optim.clip_grad_norm_(optimizer, grad_norm)
        optimizer.step()
        optimizer.zero_grad()
        parts.append(step_parts)
    if not parts:
        return {"loss": float("nan"), "steps": 0, "per_step": []}
    P = torch.stack(parts).cpu()                                         # the epoch's one synchronisation
    w = torch.tensor([task_weight, 1 - task_weight] + [1.0] * (P.shape[1] - 2))
    per_step = (P * w).sum(dim=1)
    names = ["loss_e", "loss_r", "loss_static", "loss_radius"][:P.shape[1]]
    out = {"loss": float(per_step.mean()), "steps": len(parts), "per_step": per_step.tolist()}
    out.update({n: float(P[:, i].mean()) for i, n in enumerate(names)})
    return out
