"""CPU tests of the product's host side (no GPU, no compute calls into CUDA): the C ABI exports every
symbol of include/spg_capi.h, the g2o reader / decimation / computeSubstituteEdge agree with the
oracle, and the no-device path fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sparsifyposegraph_b200 import capi
    L = capi.lib()
    hdr = open(os.path.join(ROOT, "include", "spg_capi.h")).read()
    names = set(re.findall(r"\b(spg_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 25
    for n in sorted(names):
        assert hasattr(L, n), n
    assert b"sm_100a" in L.spg_version()


def test_no_device_fails_loudly():
    import torch
    from sparsifyposegraph_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SpgError):
        capi.Context(0)


def test_record_sizes_match_python_mirror():
    from sparsifyposegraph_b200 import capi, records as R
    L = capi.lib()
    for dim in (3, 6):
        for alg in (0, 1):
            for topo in (0, 1, 3):
                for nk in (1, 2, 3, 5, 9):
                    o = capi.make_opts(topo, 1, 1.0)
                    assert L.spg_out_record_words(dim, alg, ctypes.byref(o), nk) == R.out_record_words(dim, alg, topo, 1.0, nk)


@pytest.mark.parametrize("name", ["intel", "sphere"])
def test_graph_load_matches_oracle(oracle, name):
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path(name))
    o = oracle.Graph(datasets.path(name))
    assert (g.num_vertices, g.num_edges, g.dim, g.max_vertex_id) == (o.num_vertices, o.num_edges, o.dim, o.max_vertex_id)
    assert np.array_equal(g.vertex_ids(), o.vertex_ids())
    ge, oe = g.edges(), o.edges()
    for a, b in list(zip(ge, oe))[::97]:
        assert a["uid"] == b["uid"] and list(a["v"]) == list(b["v"])
        assert np.allclose(a["meas"], b["meas"], atol=1e-15) and np.allclose(a["info"], b["info"], atol=0)


def test_decimation_matches_reference_rules(oracle):
    from sparsifyposegraph_b200 import capi
    for last, end, s in [(942, 942, 2), (941, 942, 2), (20, 20, 3), (2499, 2499, 5)]:
        assert np.array_equal(capi.decimate_global(last, end, s), oracle.decimate_global(last, end, s))
    for last in range(3, 40):
        assert np.array_equal(capi.decimate_online(last, 100, 3), oracle.decimate_online(last, 100, 3))
        assert np.array_equal(capi.decimate_cluster(last, 37, 2, 10), oracle.decimate_cluster(last, 37, 2, 10))


@pytest.mark.parametrize("name", ["intel", "sphere"])
def test_compute_substitute_edge_matches_oracle(oracle, name):
    """Online profile situation (evaluate.cpp:103-123): a new vertex links to an already marginalised one."""
    import ctypes as C
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path(name))
    o = oracle.Graph(datasets.path(name))
    L = oracle.lib()
    edges = g.edges()
    rng = np.random.default_rng(0)
    checked = 0
    for e in edges:
        a, b = int(e["v"][0]), int(e["v"][1])
        hi, lo = max(a, b), min(a, b)
        if hi - lo < 2 or lo < 4 or lo % 2 == 0:
            continue
        marg = [i for i in range(4, hi) if i % 2 != 0]   # odd ids below the new vertex are gone
        f1, t1, m1, i1 = g.compute_substitute_edge(marg, hi - 1, a, b)
        mm = np.asarray(sorted(marg), dtype=np.int32)
        f, t = C.c_int(a), C.c_int(b)
        m2, i2 = np.zeros(g.P), np.zeros(g.dim * g.dim)
        L.orc_compute_substitute_edge(o.h, mm.ctypes.data_as(C.c_void_p), len(mm), hi - 1, C.byref(f), C.byref(t),
                                      m2.ctypes.data_as(C.c_void_p), i2.ctypes.data_as(C.c_void_p))
        assert (f1, t1) == (f.value, t.value)
        assert np.allclose(m1, m2, atol=1e-12)
        assert np.allclose(i1, i2.reshape(g.dim, g.dim).T, rtol=1e-10, atol=1e-9)
        checked += 1
        if checked >= 25:
            break
    assert checked >= 5


@pytest.mark.parametrize("order,window", [("raster", 0), ("colour", 0), ("random", 0), ("random", 48), ("colour", 7)])
def test_round_scheduler_equals_sequential_removal_on_grid(oracle, order, window, monkeypatch):
    """BASELINE.json configs[4] at small scale: the wavefront rounds (product scheduler; blankets computed by the
    oracle here, no GPU) must leave exactly the graph of the one-at-a-time loop of VertexRemover::remove
    (vertex_remover.cpp:83-140), in raster order (narrow rounds), in colour order and in random order (wide rounds).
    window > 0 forces the planner's scan window (SPG_PLAN_WINDOW) far below the list length: rounds are then drawn from
    the deferred units plus a few fresh list entries — a different schedule, the same final graph. The multi-threaded
    pack / splice paths run too (SPG_HOST_THREADS=4 with thresholds far below these round widths is not needed: the
    parallel passes switch on by round width, see test below)."""
    from sparsifyposegraph_b200 import capi, synth, records as R
    if window:
        monkeypatch.setenv("SPG_PLAN_WINDOW", str(window))
    from test_gpu_graph import compare_graphs
    rows, cols = 24, 28
    data = synth.make_grid_graph(rows, cols, dim=6)
    g = synth.fill_graph(capi.Graph(dim=6), *data)
    o = synth.fill_graph(oracle.Graph(dim=6), *data)
    which = synth.grid_removal_order(rows, cols, 10, 4, order)
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    capi.rounds_begin(g, which, opts, R.ALG_NFR)
    widths = []
    while True:
        rd = capi.round_next(g)
        if rd is None:
            break
        widths.append(rd["n"])
        out = oracle.remove_round(rd["dim"], rd["algorithm"], oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rd["records"],
                                  rd["rec_off"], rd["out_off"], 0)[0]
        capi.round_apply(g, out)
    assert sum(widths) == len(which)
    assert o.marginalize(which, oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), R.ALG_NFR) == 0
    compare_graphs(g, o, tol=1e-9)
    if order != "raster" and not window:
        assert max(widths) >= 25 and len(widths) < len(which) / 8
    if window:
        assert max(widths) <= max(window, 2 * window)   # a round never exceeds leftover + fresh entries


def test_round_scheduler_parallel_passes_equal_single_thread(oracle, monkeypatch):
    """A grid wide enough for the threaded extraction / packing / splicing passes (>= 512 / 1024 / 256 units per
    thread) against the one-at-a-time oracle loop."""
    from sparsifyposegraph_b200 import capi, synth, records as R
    from test_gpu_graph import compare_graphs
    rows, cols = 150, 150
    data = synth.make_grid_graph(rows, cols, dim=6)
    g = synth.fill_graph(capi.Graph(dim=6), *data)
    o = synth.fill_graph(oracle.Graph(dim=6), *data)
    which = synth.grid_removal_order(rows, cols, 10, 4, "random")
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    capi.rounds_begin(g, which, opts, R.ALG_NFR)
    widest = 0
    while True:
        rd = capi.round_next(g)
        if rd is None:
            break
        widest = max(widest, rd["n"])
        out = oracle.remove_round(rd["dim"], rd["algorithm"], oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rd["records"],
                                  rd["rec_off"], rd["out_off"], 0)[0]
        capi.round_apply(g, out)
    assert widest >= 2048
    assert o.marginalize(which, oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), R.ALG_NFR) == 0
    compare_graphs(g, o, tol=1e-9)


@pytest.mark.parametrize("ranges", [3, 64])
def test_round_spliced_in_ranges_equals_sequential_removal(oracle, ranges, monkeypatch):
    """remove() splices every pipeline chunk of a round as its records arrive (applyBegin / applyRange / applyEnd,
    spg_remove_round_pipelined). SPG_SPLICE_RANGES cuts the splice of a CPU-driven round the same way: the final graph
    must still be the one of the one-at-a-time loop."""
    from sparsifyposegraph_b200 import capi, synth, records as R
    from test_gpu_graph import compare_graphs
    monkeypatch.setenv("SPG_SPLICE_RANGES", str(ranges))
    rows, cols = 40, 44
    data = synth.make_grid_graph(rows, cols, dim=6)
    g = synth.fill_graph(capi.Graph(dim=6), *data)
    o = synth.fill_graph(oracle.Graph(dim=6), *data)
    which = synth.grid_removal_order(rows, cols, 10, 4, "random")
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    capi.rounds_begin(g, which, opts, R.ALG_NFR)
    widest = 0
    while True:
        rd = capi.round_next(g)
        if rd is None:
            break
        widest = max(widest, rd["n"])
        out = oracle.remove_round(rd["dim"], rd["algorithm"], oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rd["records"],
                                  rd["rec_off"], rd["out_off"], 0)[0]
        capi.round_apply(g, out)
    assert widest > ranges
    assert o.marginalize(which, oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), R.ALG_NFR) == 0
    compare_graphs(g, o, tol=1e-9)
    assert g.num_edges == o.num_edges and g.num_vertices == o.num_vertices


def test_failed_blanket_in_a_ranged_splice_keeps_its_slots_dead(oracle, monkeypatch):
    """The substitutes of a round get their edge-store slots from the providers' upper bounds before any record is read:
    a blanket that comes back failed must leave its vertex and edges in place and its slots dead (not counted, not
    listed), whatever range it falls into."""
    from sparsifyposegraph_b200 import capi, synth, records as R
    monkeypatch.setenv("SPG_SPLICE_RANGES", "4")
    rows, cols = 12, 12
    data = synth.make_grid_graph(rows, cols, dim=6)
    g = synth.fill_graph(capi.Graph(dim=6), *data)
    which = synth.grid_removal_order(rows, cols, 10, 4, "random")
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    capi.rounds_begin(g, which, opts, R.ALG_NFR)
    rd = capi.round_next(g)
    assert rd["n"] >= 8
    out = oracle.remove_round(rd["dim"], rd["algorithm"], oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rd["records"],
                              rd["rec_off"], rd["out_off"], 0)[0]
    victim = 5
    nv0, ne0 = g.num_vertices, g.num_edges
    i32 = out.view(np.int32)
    new_edges = sum(int(i32[2 * int(rd["out_off"][b]) + 1]) for b in range(rd["n"]) if b != victim)
    rec = rd["records"][int(rd["rec_off"][victim]):int(rd["rec_off"][victim + 1])]
    victim_id = int(rec.view(np.int32)[2 * 4])          # first id of the record: the removed vertex
    i32[2 * int(rd["out_off"][victim])] = 1               # NOT_PD_MARGINAL
    with pytest.raises(capi.SpgError):                    # SPG_ERR_BLANKET_FAILED: reported, the other units are applied
        capi.round_apply(g, out)
    assert victim_id in list(g.vertex_ids())
    assert g.num_vertices == nv0 - (rd["n"] - 1)
    edges = g.edges()
    assert len(edges) == g.num_edges
    assert sum(1 for e in edges if e["uid"][0] >= 0) == new_edges
    assert sum(1 for e in edges if victim_id in e["v"]) >= 2      # its blanket edges are still there


def test_packed_records_have_no_unwritten_words(oracle):
    """packUnit writes every word of a record exactly once (no memset of the whole round first). With
    SPG_POISON_RECORDS=1 the staging memory is filled with 0xFF before each record is packed: the records of every round
    — pose edges, GLC factors with one (odd: padded index table) and two vertices, odd vertex and edge counts — must be
    bit-identical to the unpoisoned ones. Run in a child process: the switch is read once per process."""
    import subprocess
    import sys
    code = r'''
import hashlib, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import datasets
from oracle import pyoracle
from sparsifyposegraph_b200 import capi, records as R
h = hashlib.sha256()
for alg, topo in ((R.ALG_GLC, R.TOPO_TREE), (R.ALG_NFR, R.TOPO_TREE)):
    g = capi.Graph(datasets.path("intel"))
    for sparsity in (3, 2):          # the second pass meets the factors the first one created
        which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, sparsity)
        which = [v for v in which if v in set(g.vertex_ids().tolist())]
        capi.rounds_begin(g, which, capi.make_opts(topo, R.LIN_GLOBAL), alg)
        while True:
            rd = capi.round_next(g)
            if rd is None:
                break
            h.update(rd["records"].tobytes())
            out = pyoracle.remove_round(rd["dim"], rd["algorithm"], pyoracle.make_opts(topo, R.LIN_GLOBAL), rd["records"],
                                        rd["rec_off"], rd["out_off"], 0)[0]
            capi.round_apply(g, out)
print(h.hexdigest())
''' % (ROOT, os.path.join(ROOT, "tests"))
    digests = []
    for poison in (False, True):
        env = dict(os.environ)
        env.pop("SPG_POISON_RECORDS", None)
        if poison:
            env["SPG_POISON_RECORDS"] = "1"
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        digests.append(res.stdout.strip().splitlines()[-1])
    assert digests[0] == digests[1] and len(digests[0]) == 64
