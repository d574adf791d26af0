"""evaluate() replay loop and job-line parser (reference src/evaluate.cpp:32-221, src/main.cpp:9-103)."""
import os

import numpy as np
import pytest

import datasets
from sparsifyposegraph_b200 import records as R


def test_parse_job_lines_of_the_input_generator():
    """scripts/inputgenerator.sh:39-73 writes lines like these; defaults as in main.cpp:60-97."""
    from sparsifyposegraph_b200 import capi
    i = capi.parse_job("glc datasets/intel.g2o global tree global 2")
    assert (i.algorithm, i.profile, i.opts.topology, i.opts.lin_point, i.sparsity) == (R.ALG_GLC, 2, R.TOPO_TREE, R.LIN_GLOBAL, 2)
    assert i.kld_period == 2 ** 31 - 1 and i.use_chi2 == 0 and i.cluster_size == 100 and i.g2oname == b"datasets/intel.g2o"
    i = capi.parse_job("sen datasets/manhattan.g2o online clsubgr local 3 25 chi2 50")
    assert (i.algorithm, i.profile, i.opts.topology, i.opts.lin_point, i.sparsity) == (R.ALG_NFR, 0, 2, 0, 3)
    assert (i.kld_period, i.use_chi2, i.cluster_size) == (25, 1, 50)
    i = capi.parse_job("NONE x.g2o cluster cldense global 4 10 kld")
    assert (i.algorithm, i.profile, i.opts.topology, i.kld_period, i.use_chi2) == (-1, 1, 4, 10, 0)
    assert i.opts.chord_ratio == 1.0 and i.opts.include_intra_clique == 1
    with pytest.raises(capi.SpgError):
        capi.parse_job("glc")


def _head_of(name, nverts):
    """The first `nverts` vertices of a dataset as a new product graph (ids stay contiguous from 0)."""
    from sparsifyposegraph_b200 import capi
    full = capi.Graph(datasets.path(name))
    g = capi.Graph(dim=full.dim)
    for vid in range(nverts):
        g.add_vertex(vid, full.vertex_pose(vid))
    for e in full.edges():
        if max(e["v"]) < nverts:
            g.add_edge(int(e["v"][0]), int(e["v"][1]), e["meas"], e["info"])
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ["glc", "se2"])
def test_global_profile_equals_the_manual_pipeline(alg, tmp_path):
    """global profile: nothing happens until the last vertex; then optimise, marginalise every 2nd vertex, optimise, KLD —
    the same numbers as calling the pieces by hand, and the reference's result-file layout."""
    from sparsifyposegraph_b200 import capi
    ctx = capi.Context(0)
    gw = _head_of("intel", 300)
    info = capi.parse_job(f"{alg} datasets/intel.g2o global tree global 2")
    res = capi.evaluate(ctx, gw, info, destdir=str(tmp_path), want_graphs=True)
    assert res["n_samples"] == 1 and res["samples"][0][0] == 299 and res["n_marginalize_calls"] == 1
    which = capi.decimate_global(299, 299, 2)
    assert res["n_marginalized"] == len(which)
    # by hand
    base, inc = _head_of("intel", 300), _head_of("intel", 300)
    base.optimize(ctx)
    inc.optimize(ctx)
    algorithm = R.ALG_GLC if alg == "glc" else R.ALG_NFR
    inc.marginalize(ctx, which, capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), algorithm)
    inc.optimize(ctx)
    kld, _ = base.kld(ctx, inc)
    assert kld > 0 and abs(res["last_value"] - kld) <= 1e-4 * kld
    assert np.array_equal(res["incremental"].vertex_ids(), inc.vertex_ids())
    assert res["marginal_nodes"] == len(inc.vertex_ids()) - 1 and res["baseline_nodes"] == 299
    stem = tmp_path / "global" / "2" / "intel" / f"{alg}_tree_g"
    lines = open(str(stem) + ".kld").read().split()
    assert lines[0] == "299" and abs(float(lines[1]) - res["last_value"]) <= 1e-5 * kld
    txt = open(str(stem) + ".txt").read()
    assert txt.startswith(("GLC" if alg == "glc" else "SE2") + " Tree") and "baseline:     nodes = 299; edges = " in txt and "last kld: " in txt
    ctx.close()


@pytest.mark.gpu
def test_online_profile_uses_substitute_edges_and_samples_periodically():
    """online profile, sparsity 3: a vertex is marginalised as the replay goes, later vertices that link to a removed one
    get a computeSubstituteEdge edge; KLD sampled every 40 vertices and at the end, non-negative and finite."""
    from sparsifyposegraph_b200 import capi
    ctx = capi.Context(0)
    gw = _head_of("intel", 160)
    info = capi.parse_job("sen datasets/intel.g2o online tree global 3 40")
    res = capi.evaluate(ctx, gw, info, want_graphs=True)
    assert [v for v, _ in res["samples"]] == [40, 80, 120, 159]
    assert all(np.isfinite(k) and k > -1e-6 for _, k in res["samples"])
    assert res["n_marginalized"] == sum(len(capi.decimate_online(i, 159, 3)) for i in range(4, 160))
    kept = set(res["incremental"].vertex_ids().tolist())
    assert kept == set(range(160)) - {v for i in range(4, 160) for v in capi.decimate_online(i, 159, 3).tolist()}
    assert res["baseline_nodes"] == 159 and res["marginal_nodes"] == len(kept) - 1
    # delta chi2 flavour of the same job
    info2 = capi.parse_job("sen datasets/intel.g2o online tree global 3 80 chi2")
    res2 = capi.evaluate(ctx, gw, info2)
    assert [v for v, _ in res2["samples"]] == [80, 159] and all(np.isfinite(k) for _, k in res2["samples"])
    ctx.close()
