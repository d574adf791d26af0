"""Correlated topologies on the device: CliqueySubgraph and CliqueyDense group the Chow-Liu
spanning tree into cliques (PseudoChowLiu::fillCliques, reference src/pseudo_chow_liu.cpp:198-251) and emit one
MultiEdgeCorrelated per clique with the closed-form information X_c = (J_c Sigma J_c^T)^-1 (src/logdet_function.cpp:236-279);
later blankets contain those multi-edges as INPUT factors (src/multi_edge_correlated.hpp:96-140). CUDA path through the
C ABI against the oracle: clique structure bit-exact, informations <= 1e-9 relative Frobenius, projected KLD <= 1e-6."""
import numpy as np
import pytest

import datasets
from sparsifyposegraph_b200 import records as R, synth
from test_gpu_parity import run_both, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sparsifyposegraph_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("topology", [R.TOPO_CLIQUEY_SUBGRAPH, R.TOPO_CLIQUEY_DENSE], ids=["clsubgr", "cldense"])
@pytest.mark.parametrize("dim,n,variant,B,chord", [
    (6, 3, "ring", 64, 1.0), (6, 4, "ring", 128, 1.0), (6, 5, "ring", 256, 1.0), (6, 6, "star", 128, 1.0), (6, 8, "ring", 96, 1.0),
    (6, 8, "ring", 64, 0.5), (6, 12, "ring", 32, 1.0), (6, 16, "ring", 16, 1.0), (6, 16, "ring", 8, 2.5),
    (3, 4, "ring", 64, 1.0), (3, 7, "ring", 128, 1.0), (3, 14, "star", 32, 1.0), (3, 24, "ring", 8, 1.5),
])
@pytest.mark.parametrize("flags", [0, 1], ids=["gauge-shortcut", "force-eigen"])
def test_cliquey_blanket_parity(ctx, oracle, topology, dim, n, variant, B, chord, flags):
    blk = synth.make_blankets(n, B, dim=dim, variant=variant, seed=4000 + 10 * n + dim)
    out_off, (ro, _, rt, rw), (go, gt, gw) = run_both(ctx, oracle, blk, R.ALG_NFR, topology, chord_ratio=chord, flags=flags)
    worst_x = worst_kld = 0.0
    n_multi = n_entries = 0
    for b in range(B):
        r = R.parse_out(ro, out_off, b, dim, R.ALG_NFR, topology, n - 1)
        g = R.parse_out(go, out_off, b, dim, R.ALG_NFR, topology, n - 1)
        assert g["status"] == r["status"] == 0, (b, g["status"], r["status"])
        assert g["n_edges"] == r["n_edges"], (b, g["n_edges"], r["n_edges"])
        assert sum(e["nmeas"] for e in g["edges"]) == n - 2          # the cliques partition the spanning tree
        for eg, er in zip(g["edges"], r["edges"]):
            assert eg["nmeas"] == er["nmeas"] and eg["pairs"] == er["pairs"], (b, eg["pairs"], er["pairs"])   # bit-exact structure
            assert np.allclose(eg["meas"], er["meas"], atol=1e-12)
            worst_x = max(worst_x, rel(eg["info"], er["info"]))
            n_multi += eg["nmeas"] > 1
            n_entries += 1
        if flags == 0:
            assert not (g["flags"] & 8)
            worst_kld = max(worst_kld, abs(g["kld"] - r["kld"]) / max(abs(r["kld"]), 1e-3))   # KLD is exactly 0 when the tree is the whole graph
    print(f"dim {dim} n {n} topo {topology} chord {chord}: {n_entries} entries ({n_multi} correlated), worst X {worst_x:.2e}, kld {worst_kld:.2e}")
    assert worst_x <= 1e-9, worst_x
    assert worst_kld <= 1e-6, worst_kld
    if n >= 4:
        assert n_multi > 0
    if topology == R.TOPO_CLIQUEY_DENSE:
        assert n_entries == B


@pytest.mark.parametrize("name,topology,sparsity,count", [
    ("intel", R.TOPO_CLIQUEY_SUBGRAPH, 2, 469), ("intel", R.TOPO_CLIQUEY_DENSE, 3, 200), ("sphere", R.TOPO_CLIQUEY_SUBGRAPH, 2, 400),
    ("manhattan", R.TOPO_CLIQUEY_SUBGRAPH, 2, 600),
])
def test_cliquey_graph_level_matches_sequential_oracle(ctx, oracle, name, topology, sparsity, count):
    """Job lines `sen <dataset> global clsubgr|cldense global <s>` (scripts/inputgenerator.sh:39-73): multi-edges created in
    one round are input factors of the blankets of later rounds."""
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path(name))
    o = oracle.Graph(datasets.path(name))
    which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, sparsity)[:count]
    st = g.marginalize(ctx, which, capi.make_opts(topology, R.LIN_GLOBAL), R.ALG_NFR)
    assert st["n_failed"] == 0 and st["n_blankets"] > 0
    assert o.marginalize(which, oracle.make_opts(topology, R.LIN_GLOBAL), R.ALG_NFR) == 0
    ge, oe = g.edges(), o.edges()
    assert np.array_equal(g.vertex_ids(), o.vertex_ids()) and len(ge) == len(oe)
    worst, n_multi = 0.0, 0
    for a, b in zip(ge, oe):
        assert a["uid"] == b["uid"] and a["kind"] == b["kind"] and list(a["v"]) == list(b["v"]), (a["uid"], a["v"], b["v"])
        n_multi += a["kind"] == 2   # vertex list order (order of appearance of the measurements' vertices) and rows compared above / below
        assert a["rows"] == b["rows"]
        assert np.allclose(np.asarray(a["meas"]).reshape(-1), np.asarray(b["meas"]).reshape(-1), atol=1e-11)
        worst = max(worst, rel(np.asarray(a["info"]), np.asarray(b["info"])))
    print(f"{name} topo {topology}: {st['n_blankets']} blankets in {st['n_rounds']} rounds, {n_multi} multi-edges left, worst X {worst:.2e}")
    assert n_multi > 0 and worst <= 1e-9, worst


def test_multi_edges_round_trip_through_g2o_text(ctx, tmp_path):
    """f4: a CliqueySubgraph-sparsified graph saved with MULTI_EDGE_* factors and loaded back is the same graph."""
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path("intel"))
    which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, 2)[:150]
    g.marginalize(ctx, which, capi.make_opts(R.TOPO_CLIQUEY_SUBGRAPH, R.LIN_GLOBAL), R.ALG_NFR)
    path = str(tmp_path / "cl.g2o")
    g.save(path)
    assert "MULTI_EDGE_SE2_ISAM" in open(path).read()
    h = capi.Graph(path)
    ea, eb = g.edges(), h.edges()
    assert len(ea) == len(eb)
    for x, y in zip(ea, eb):
        assert x["kind"] == y["kind"] and list(x["v"]) == list(y["v"]) and np.array_equal(x["info"], y["info"])
        if x["kind"] == 2:
            assert np.array_equal(x["pairs"], y["pairs"])


def test_kld_and_optimiser_handle_multi_edges(ctx, oracle):
    """The evaluator / optimiser assemble MultiEdgeCorrelated factors: KLD of a CliqueySubgraph-sparsified intel against the
    CPU harness, a lower KLD than the plain tree (the cliques carry the correlations the tree drops), and the replay loop on
    a `clsubgr` job line."""
    import kld_harness as K
    from sparsifyposegraph_b200 import capi
    full = capi.Graph(datasets.path("intel"))
    klds = {}
    for topo in (R.TOPO_TREE, R.TOPO_CLIQUEY_SUBGRAPH):
        g = capi.Graph(datasets.path("intel"))
        which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, 2)
        g.marginalize(ctx, which, capi.make_opts(topo, R.LIN_GLOBAL), R.ALG_NFR)
        klds[topo], _ = full.kld(ctx, g)
        if topo == R.TOPO_CLIQUEY_SUBGRAPH:
            assert any(e["kind"] == 2 for e in g.edges())
            ref, _ = K.full_graph_kld(oracle, 3, K.poses_of(full), full.edges(), K.poses_of(g), g.edges())
            assert abs(klds[topo] - ref) <= 1e-6 * abs(ref)
            chi0 = g.chi2(ctx)
            st = g.optimize(ctx)
            assert st["chi2_final"] <= chi0 * (1 + 1e-12)
    print("intel KLD tree", klds[R.TOPO_TREE], "cliquey subgraph", klds[R.TOPO_CLIQUEY_SUBGRAPH])
    assert 0 < klds[R.TOPO_CLIQUEY_SUBGRAPH] < klds[R.TOPO_TREE]
    # evaluate() on a correlated topology
    small = capi.Graph(dim=3)
    for vid in range(200):
        small.add_vertex(vid, full.vertex_pose(vid))
    for e in full.edges():
        if max(e["v"]) < 200:
            small.add_edge(int(e["v"][0]), int(e["v"][1]), e["meas"], e["info"])
    res = capi.evaluate(ctx, small, capi.parse_job("sen datasets/intel.g2o global clsubgr global 2"))
    # the first 200 poses of intel are almost a chain: the cliques keep every correlation and the KLD is 0 up to rounding
    assert res["n_samples"] == 1 and np.isfinite(res["last_value"]) and -1e-9 < res["last_value"] < 1.0 and res["n_marginalized"] == 98
