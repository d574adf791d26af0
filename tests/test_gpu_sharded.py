"""Sharded C-ABI entry points on one GPU (communicator of one rank: no NCCL traffic, same code path for the shard
plan, the pipeline steps and the output placement). The multi-rank run is tools/sharded_check.py under torchrun
(needs >= 2 GPUs; results in profiles/), the multi-rank host logic tests/test_distributed.py (gloo, CPU)."""
import numpy as np
import pytest

from sparsifyposegraph_b200 import records as R, synth

pytestmark = pytest.mark.gpu


def _round(sizes, per, seed=7):
    blks = [synth.make_blankets(n, per, dim=6, variant="ring", seed=seed + n) for n in sizes]
    rec = np.concatenate([b["records"] for b in blks])
    rec_off = np.zeros(1, dtype=np.int64)
    for b in blks:
        rec_off = np.concatenate([rec_off, b["rec_off"][1:] + rec_off[-1]])
    out_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, R.n_kept_of(rec, rec_off))
    return rec, rec_off, out_off


def test_sharded_round_single_rank_equals_plain_round(monkeypatch):
    from sparsifyposegraph_b200 import capi
    monkeypatch.setenv("SPG_CHUNK_BYTES", str(1 << 20))
    ctx = capi.Context(0)
    ctx.comm_init(1, 0)
    assert ctx.nranks == 1 and ctx.rank == 0
    rec, rec_off, out_off = _round((2, 3, 5, 9, 16), 400)
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL, 1.0, flags=2)
    ref, _, _ = ctx.remove_round(6, R.ALG_NFR, opts, rec, rec_off, out_off)
    for root in (-1, 0):
        got, info = ctx.remove_round_sharded(6, R.ALG_NFR, opts, rec, rec_off, out_off, root=root)
        assert np.array_equal(got, ref)
        assert info["n_blankets_mine"] == len(rec_off) - 1 and info["steps"] >= 2
        assert info["d2h_bytes"] == int(out_off[-1]) * 8
        assert info["gather_bytes"] == 0
    ctx.close()


def test_sharded_device_round_single_rank():
    import torch
    from sparsifyposegraph_b200 import capi
    ctx = capi.Context(0)
    ctx.comm_init(1, 0)
    blk = synth.make_blankets(6, 500, dim=6, variant="ring", seed=11)
    B = blk["B"]
    out_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, np.full(B, 5))
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL, 1.0, flags=2)
    ref, _, _ = ctx.remove_round(6, R.ALG_NFR, opts, blk["records"], blk["rec_off"], out_off)
    d_rec = torch.from_numpy(blk["records"].view(np.int64)).cuda()
    d_ro, d_oo = torch.from_numpy(blk["rec_off"]).cuda(), torch.from_numpy(out_off).cuda()
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.int64, device="cuda")
    bounds = np.array([0, B], dtype=np.int32)
    ctx.remove_round_sharded_device(6, R.ALG_NFR, opts, B, d_rec.data_ptr(), d_ro.data_ptr(), d_oo.data_ptr(),
                                    d_out.data_ptr(), bounds, out_off[bounds], 6, blk["E"], root=-1)
    ctx.comm_join()
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), ref)
    ctx.close()


def test_sharded_needs_communicator_for_bad_root():
    from sparsifyposegraph_b200 import capi
    ctx = capi.Context(0)
    rec, rec_off, out_off = _round((3,), 8)
    with pytest.raises(capi.SpgError):
        ctx.remove_round_sharded(6, R.ALG_NFR, capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rec, rec_off, out_off, root=3)
    ctx.close()
