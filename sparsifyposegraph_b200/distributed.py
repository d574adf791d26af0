"""Multi-GPU node removal: the blankets of a wavefront round are independent, so every rank takes a
contiguous shard of the round, runs it on its own GPU, and the substitute-edge records are gathered
back (the only exchange of the path: torch.distributed all_gather — NCCL over NVLink on GPUs, gloo in
the CPU tests). Every rank holds the same host-side graph and applies the same gathered round, so the
graphs stay identical without any other communication.
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_bounds(rec_off, world):
    """Contiguous split of a round's blankets into `world` shards balanced by record size cubed (a proxy
    for the per-blanket cost, which is cubic in the blanket dimension). Returns world+1 boundaries."""
    n = len(rec_off) - 1
    if n == 0:
        return [0] * (world + 1)
    w = np.diff(np.asarray(rec_off, dtype=np.float64)) ** 1.5
    c = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(c, c[-1] * r / world)))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def marginalize_sharded(graph, which, opts, algorithm, compute, rank=0, world=1, group=None, device=None):
    """VertexRemover::remove(which) with every round sharded over `world` ranks.

    compute(dim, algorithm, opts, records, rec_off, out_off) -> uint64 output buffer of a sub-round
    (in production `lambda *a: ctx.remove_round(*a)[0]` on this rank's GPU).
    Returns the number of rounds."""
    capi.rounds_begin(graph, which, opts, algorithm)
    rounds = 0
    while True:
        rd = capi.round_next(graph)
        if rd is None:
            break
        bounds = shard_bounds(rd["rec_off"], world)
        b0, b1 = bounds[rank], bounds[rank + 1]
        ro, oo = rd["rec_off"], rd["out_off"]
        if b1 > b0:
            local = compute(rd["dim"], rd["algorithm"], rd["opts"], rd["records"][ro[b0]:ro[b1]], ro[b0:b1 + 1] - ro[b0],
                            oo[b0:b1 + 1] - oo[b0])
            local = np.ascontiguousarray(local, dtype=np.uint64)
        else:
            local = np.zeros(0, dtype=np.uint64)
        if world > 1:
            out = gather_outputs(local, [int(oo[bounds[r + 1]] - oo[bounds[r]]) for r in range(world)], group, device)
        else:
            out = local
        capi.round_apply(graph, out)
        rounds += 1
    return rounds


def gather_outputs(local, sizes, group=None, device=None):
    """all_gather of the per-rank output slices (padded to the largest) -> concatenated full round output."""
    import torch
    import torch.distributed as dist
    pad = max(max(sizes), 1)
    t = torch.zeros(pad, dtype=torch.int64, device=device)
    if len(local):
        t[:len(local)] = torch.from_numpy(local.view(np.int64)).to(t.device)
    world = len(sizes)
    full = torch.empty(world * pad, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(full, t, group=group) if t.is_cuda else dist.all_gather(list(full.view(world, pad).unbind(0)), t, group=group)
    full = full.view(world, pad).cpu().numpy().view(np.uint64)
    return np.concatenate([full[r, :sizes[r]] for r in range(world)])
