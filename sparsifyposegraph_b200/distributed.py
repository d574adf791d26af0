"""Multi-GPU node removal: the blankets of a wavefront round are independent, so every rank takes a
contiguous shard of the round, runs it on its own GPU, and the substitute-edge records are gathered
back. The product path is C++ behind the C ABI (csrc/spg_comm.cu: spg_comm_init, spg_remove_round_sharded, NCCL
directly on the device output buffers); this module only (1) hands the NCCL unique id around with
torch.distributed and (2) keeps a Python restatement of the round loop with a pluggable per-shard engine and a gloo
gather, so that the sharding / gather / splice logic is tested on CPU (tests/test_distributed.py).
Every rank holds the same host-side graph and applies the same gathered round, so the graphs stay identical
without any other communication.
"""
from __future__ import annotations

import numpy as np

from . import capi


def init_comm(ctx, rank, world, group=None):
    """Give `ctx` its NCCL communicator (spg_comm_init): rank 0 draws the unique id through the library and it is
    shipped over the already-initialised torch.distributed group (any transport would do: it is 128 bytes)."""
    import torch.distributed as dist
    box = [capi.comm_unique_id() if rank == 0 else None] if world > 1 else [None]
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(world, rank, box[0])


def shard_bounds(rd, world):
    """spg_shard_bounds of a round dict (capi.round_next): contiguous shards balanced by the per-blanket cost
    model of the library. Returns world+1 boundaries."""
    if rd["n"] == 0:
        return [0] * (world + 1)
    return [int(x) for x in capi.shard_bounds(rd["dim"], rd["algorithm"], rd["opts"], rd["records"], rd["rec_off"],
                                              rd["out_off"], world)]


def marginalize_sharded(graph, which, opts, algorithm, compute, rank=0, world=1, group=None, device=None):
    """VertexRemover::remove(which) with every round sharded over `world` ranks.

    compute(dim, algorithm, opts, records, rec_off, out_off) -> uint64 output buffer of a sub-round
    (this Python loop exists for the CPU tests, where the oracle stands in for the GPU; on GPUs the whole loop is
    C++: `distributed.init_comm(ctx, rank, world)` once, then `graph.marginalize(ctx, ...)` is collective and every
    round is one spg_remove_round_sharded with the NCCL gather on the device buffers).
    Returns the number of rounds."""
    capi.rounds_begin(graph, which, opts, algorithm)
    rounds = 0
    while True:
        rd = capi.round_next(graph)
        if rd is None:
            break
        bounds = shard_bounds(rd, world)
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
