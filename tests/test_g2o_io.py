"""g2o text I/O of sparsified graphs: GLC_EDGE / GLC_REPARAM_*, MULTI_EDGE_*, EDGE_SE2_ISAM
(reference src/glc_edge.cpp:64-118, src/multi_edge_correlated.hpp:183-267, src/edge_types.cpp:68-76). CPU only: the GLC
factors come from the round planner with the oracle as the per-blanket engine."""
import os

import numpy as np
import pytest

import datasets
from sparsifyposegraph_b200 import records as R


def _glc_graph(oracle, name, topo):
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path(name))
    which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, 2)[:120]
    opts = capi.make_opts(topo, R.LIN_GLOBAL)
    capi.rounds_begin(g, which, opts, R.ALG_GLC)
    while True:
        rd = capi.round_next(g)
        if rd is None:
            break
        out = oracle.remove_round(rd["dim"], rd["algorithm"], oracle.make_opts(topo, R.LIN_GLOBAL), rd["records"], rd["rec_off"],
                                  rd["out_off"], 0)[0]
        capi.round_apply(g, out)
    return g


def _same(a, b):
    assert np.array_equal(a.vertex_ids(), b.vertex_ids())
    for vid in a.vertex_ids()[:50]:
        assert np.allclose(a.vertex_pose(vid), b.vertex_pose(vid), rtol=0, atol=2e-15)   # quaternions are re-normalised on load
    ea, eb = a.edges(), b.edges()
    assert len(ea) == len(eb)
    for x, y in zip(ea, eb):
        assert x["kind"] == y["kind"] and x["rows"] == y["rows"] and list(x["v"]) == list(y["v"])
        assert np.allclose(x["meas"], y["meas"], rtol=0, atol=2e-15) and np.array_equal(x["info"], y["info"])   # 17 digits: bit-exact
        if x["kind"] == 2:
            assert np.array_equal(x["pairs"], y["pairs"])


@pytest.mark.parametrize("name,topo", [("intel", R.TOPO_TREE), ("sphere", R.TOPO_TREE), ("intel", R.TOPO_DENSE)])
def test_glc_graph_round_trips_through_g2o_text(oracle, tmp_path, name, topo):
    from sparsifyposegraph_b200 import capi
    g = _glc_graph(oracle, name, topo)
    kinds = {e["kind"] for e in g.edges()}
    assert kinds == {0, 1}
    path = str(tmp_path / "sparse.g2o")
    g.save(path)
    text = open(path).read()
    assert "GLC_EDGE" in text and " || GLC_REPARAM_" in text
    assert ("EDGE_SE2_ISAM" in text) == (g.dim == 3)
    _same(g, capi.Graph(path))
    # the documented line layout (test_marginalize_within_window.cpp:198-206): ids || reparam rows cols meas W info
    line = next(l for l in text.splitlines() if l.startswith("GLC_EDGE"))
    ids, rest = line[len("GLC_EDGE"):].split("||")
    tok = rest.split()
    rows, cols = int(tok[1]), int(tok[2])
    assert cols == g.dim * len(ids.split()) and len(tok) == 3 + cols + rows * cols + rows * (rows + 1) // 2


@pytest.mark.parametrize("dim", [3, 6])
def test_multi_edge_round_trip_and_reference_format(tmp_path, dim):
    from sparsifyposegraph_b200 import capi, synth
    rng = np.random.default_rng(dim)
    P = 3 if dim == 3 else 7
    g = capi.Graph(dim=dim)
    poses = synth.make_grid_graph(2, 3, dim=dim)[0]
    for i, p in enumerate(poses):
        g.add_vertex(10 + i, p)
    g.add_edge(10, 11, poses[1], np.eye(dim))
    nm = 3
    A = rng.normal(size=(dim * nm, dim * nm))
    info = A @ A.T + np.eye(dim * nm)
    meas = np.concatenate([poses[2], poses[3], poses[4]])
    g.add_factor(2, [12, 13, 15, 14], dim * nm, meas, info, pairs=[0, 1, 1, 2, 1, 3])
    path = str(tmp_path / "multi.g2o")
    g.save(path)
    text = open(path).read().splitlines()
    k = next(i for i, l in enumerate(text) if l.startswith("MULTI_EDGE_"))
    assert text[k - 1].split() == ["#SPG_MULTI_PAIRS", "0", "1", "1", "2", "1", "3"]
    head, rest = text[k].split("||")
    assert head.split()[1:] == ["12", "13", "15", "14"]
    tok = rest.split()
    n = dim * nm
    assert tok[:2] == [str(nm), str(P)] and len(tok) == 2 + nm * P + n * (n + 1) // 2     # multi_edge_correlated.hpp:226-267
    _same(g, capi.Graph(path))
    # without the pairs line the mapping is unrecoverable (as in the reference): a clear error, not a guess
    open(path, "w").write("\n".join(l for l in text if not l.startswith("#SPG")) + "\n")
    with pytest.raises(capi.SpgError):
        capi.Graph(path)


def test_reader_accepts_edge_before_vertex_and_skips_fix(tmp_path):
    from sparsifyposegraph_b200 import capi
    path = str(tmp_path / "g.g2o")
    open(path, "w").write("EDGE_SE2 0 1 1 0 0.1 1 0 0 1 0 1\nVERTEX_SE2 0 0 0 0\nFIX 0\nVERTEX_SE2 1 1 0 0.1\n# comment\n")
    g = capi.Graph(path)
    assert g.num_vertices == 2 and g.num_edges == 1
