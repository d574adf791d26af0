"""N>1 host path on CPU (gloo, world_size 2): rounds sharded over ranks, output records gathered, every
rank applies the same round. The per-blanket arithmetic is done by the ORACLE here (no GPU in this
container); what is under test is the product's round planner / sharding / gather / splice."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_compute(dim, algorithm, opts, records, rec_off, out_off):
    from oracle import pyoracle
    o = pyoracle.make_opts(opts.topology, opts.lin_point, opts.chord_ratio, bool(opts.include_intra_clique))
    return pyoracle.remove_round(dim, algorithm, o, records, rec_off, out_off, 1)[0]


def _edges_signature(g):
    sig = []
    for e in g.edges():
        sig.append((e["uid"], tuple(int(v) for v in e["v"]), np.round(e["info"], 6).tobytes()))
    return sig


def _worker(rank, world, port, path, which, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparsifyposegraph_b200 import capi, distributed, records as R
    g = capi.Graph(path)
    rounds = distributed.marginalize_sharded(g, which, capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), R.ALG_NFR,
                                             _oracle_compute, rank=rank, world=world)
    ids = g.vertex_ids().tolist()
    infos = np.concatenate([e["info"].reshape(-1) for e in g.edges()])
    pairs = [tuple(int(v) for v in e["v"]) for e in g.edges()]
    q.put((rank, rounds, ids, pairs, infos))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_and_balance():
    """spg_shard_bounds (C++, pure host): contiguous, covering, balanced by the cost model."""
    from sparsifyposegraph_b200 import capi, records as R, synth
    blks = [synth.make_blankets(n, 200, dim=6, variant="ring", seed=n) for n in (3, 5, 9, 16)]
    rec = np.concatenate([b["records"] for b in blks])
    rec_off = np.zeros(1, dtype=np.int64)
    for b in blks:
        rec_off = np.concatenate([rec_off, b["rec_off"][1:] + rec_off[-1]])
    nk = R.n_kept_of(rec, rec_off)
    out_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, nk)
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    nb = len(rec_off) - 1
    nvs = [int(rec[o:o + 1].view(np.int32)[0]) for o in rec_off[:-1]]
    cost = np.array([1.0 if n <= 2 else (1.6 if n == 3 else 3.4 * (n - 2.6)) for n in nvs])   # NFR tree model, spg_comm.cu
    for world in (1, 2, 4, 8):
        b = capi.shard_bounds(6, R.ALG_NFR, opts, rec, rec_off, out_off, world)
        assert b[0] == 0 and b[-1] == nb and all(x <= y for x, y in zip(b, b[1:]))
        loads = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert loads.max() <= loads.mean() * 1.02 + cost.max()
    # fewer blankets than ranks: empty shards allowed
    b = capi.shard_bounds(6, R.ALG_NFR, opts, rec[:rec_off[2]], rec_off[:3], out_off[:3], 8)
    assert b[0] == 0 and b[-1] == 2 and all(x <= y for x, y in zip(b, b[1:]))


def test_two_rank_sharded_removal_matches_sequential_oracle(oracle):
    path = datasets.path("intel")
    which = [i for i in range(4, 943) if i % 2]
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, which, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle.Graph(path)
    from sparsifyposegraph_b200 import records as R
    assert o.marginalize(which, oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), R.ALG_NFR) == 0
    oe = o.edges()
    ref_pairs = [tuple(int(v) for v in e["v"]) for e in oe]
    ref_info = np.concatenate([e["info"].reshape(-1) for e in oe])
    for rank, rounds, ids, pairs, infos in res:
        assert rounds < 40
        assert ids == o.vertex_ids().tolist()
        assert pairs == ref_pairs
        assert np.linalg.norm(infos - ref_info) <= 1e-12 * np.linalg.norm(ref_info)
    assert np.array_equal(res[0][4], res[1][4])    # both ranks hold the same graph
