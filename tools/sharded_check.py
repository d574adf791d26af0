#!/usr/bin/env python
"""Multi-GPU correctness check of the sharded C-ABI path (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_check.py

1. spg_remove_round_sharded (root = -1 and root = 0) on a mixed-size SE3 round == spg_remove_round on one GPU, bit for bit.
2. spg_remove_round_sharded_device + spg_comm_join: every rank's device buffer ends up complete and identical.
3. spg_graph_marginalize with a communicator on the context (collective) == the single-rank result, edge for edge.
Prints one JSON line from rank 0; exit code 1 on any mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from sparsifyposegraph_b200 import capi, distributed, records as R, synth
    import datasets

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["SPG_CHUNK_BYTES"] = str(4 << 20)  # read at spg_create: several pipeline steps per round
    ctx = capi.Context(local)
    distributed.init_comm(ctx, rank, world)
    solo = capi.Context(local)  # no communicator: the single-GPU answer
    res = {"world": world, "nccl": capi.lib().spg_comm_nccl_version()}
    ok = True

    # ---- 1. host-buffer round --------------------------------------------------------------------------------
    blks = [synth.make_blankets(n, 3000, dim=6, variant="ring", seed=100 + n) for n in (2, 3, 5, 8, 12, 16)]
    rec = np.concatenate([b["records"] for b in blks])
    rec_off = np.zeros(1, dtype=np.int64)
    for b in blks:
        rec_off = np.concatenate([rec_off, b["rec_off"][1:] + rec_off[-1]])
    nk = R.n_kept_of(rec, rec_off)
    out_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, nk)
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL, 1.0, flags=2)
    ref, _, _ = solo.remove_round(6, R.ALG_NFR, opts, rec, rec_off, out_off)
    for root in (-1, 0):
        got, info = ctx.remove_round_sharded(6, R.ALG_NFR, opts, rec, rec_off, out_off, root=root)
        if root < 0 or rank == root:
            same = bool(np.array_equal(got, ref))
        else:  # a non-root rank holds the records of its own blankets only: every non-zero record must be the right one
            nz = [b for b in range(len(out_off) - 1) if got[out_off[b]:out_off[b + 1]].any()]
            same = len(nz) > 0 and len(nz) <= info["n_blankets_mine"] and all(
                np.array_equal(got[out_off[b]:out_off[b + 1]], ref[out_off[b]:out_off[b + 1]]) for b in nz)
        ok &= same
        res[f"host_round_root{root}"] = {"same": same, **info} if rank == 0 else None

    # ---- 2. device-resident round --------------------------------------------------------------------------
    blk = blks[3]
    B = blk["B"]
    o_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, np.full(B, blk["n"] - 1))
    ref_d, _, _ = solo.remove_round(6, R.ALG_NFR, opts, blk["records"], blk["rec_off"], o_off)
    d_rec = torch.from_numpy(blk["records"].view(np.int64)).cuda()
    d_ro = torch.from_numpy(blk["rec_off"]).cuda()
    d_oo = torch.from_numpy(o_off).cuda()
    d_out = torch.zeros(int(o_off[-1]), dtype=torch.int64, device="cuda")
    bounds = np.array([B * r // world for r in range(world + 1)], dtype=np.int32)
    ctx.remove_round_sharded_device(6, R.ALG_NFR, opts, B, d_rec.data_ptr(), d_ro.data_ptr(), d_oo.data_ptr(), d_out.data_ptr(),
                                    bounds, o_off[bounds], blk["n"], blk["E"], root=-1)
    ctx.comm_join()
    ctx.sync()
    same = bool(np.array_equal(d_out.cpu().numpy().view(np.uint64), ref_d))
    ok &= same
    res["device_round"] = same

    # ---- 3. graph level, collective ----------------------------------------------------------------------------
    for name, alg, topo in (("intel", R.ALG_GLC, R.TOPO_TREE), ("sphere", R.ALG_NFR, R.TOPO_TREE)):
        g1, g2 = capi.Graph(datasets.path(name)), capi.Graph(datasets.path(name))
        which = capi.decimate_global(g1.max_vertex_id, g1.max_vertex_id, 2)
        o = capi.make_opts(topo, R.LIN_GLOBAL)
        g1.marginalize(solo, which, o, alg)
        st = g2.marginalize(ctx, which, o, alg)
        e1, e2 = g1.edges(), g2.edges()
        same = len(e1) == len(e2) and all(np.array_equal(a["v"], b["v"]) and np.array_equal(a["info"], b["info"]) and
                                         np.array_equal(a["meas"], b["meas"]) for a, b in zip(e1, e2))
        ok &= same
        res[f"graph_{name}"] = {"same": same, "rounds": st["n_rounds"], "edges": len(e2)}

    if world > 1:
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    res["ok"] = ok
    if rank == 0:
        print(json.dumps(res))
    ctx.close()
    solo.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
