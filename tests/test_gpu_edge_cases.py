"""GPU edge cases (-m gpu): the situations the reference's data and code paths contain beyond the
regular sweep — empty rounds, parallel (duplicate) edges, several removed vertices in one blanket
(Dense / CliqueyDense extended blankets), weakly
constrained blankets (chooseDimensions branch of logdet_function.cpp:42-60) and non-PD marginals."""
import numpy as np
import pytest

from sparsifyposegraph_b200 import records as R
from sparsifyposegraph_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sparsifyposegraph_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def both(ctx, oracle, dim, recs, nks, algorithm, topology, flags=0):
    from sparsifyposegraph_b200 import capi
    records, rec_off = R.concat_records(recs)
    out_off = R.out_offsets(dim, algorithm, topology, 1.0, nks)
    ro = oracle.remove_round(dim, algorithm, oracle.make_opts(topology, 1), records, rec_off, out_off, 0)[0]
    go = ctx.remove_round(dim, algorithm, capi.make_opts(topology, 1, flags=flags), records, rec_off, out_off)[0]
    res = []
    for b, nk in enumerate(nks):
        res.append((R.parse_out(go, out_off, b, dim, algorithm, topology, nk),
                    R.parse_out(ro, out_off, b, dim, algorithm, topology, nk)))
    return res


def se3_edge(rng, poses, i, j, info=None, noise=(0.05, 0.02)):
    z = synth.se3_compose(synth.se3_compose(synth.se3_inverse(poses[i]), poses[j]),
                          synth.se3_exp_small(rng, (), noise[0], noise[1]))
    return {"kind": 0, "v": [i, j], "meas": z, "info": synth.random_info(rng, (), 6) if info is None else info}


def test_empty_round(ctx):
    from sparsifyposegraph_b200 import capi
    out, _, _ = ctx.remove_round(6, R.ALG_NFR, capi.make_opts(0, 1), np.zeros(0, np.uint64), np.zeros(1, np.int64),
                                 np.zeros(1, np.int64))
    assert len(out) == 0


def test_parallel_duplicate_edges(ctx, oracle):
    """manhattan / intel contain duplicated vertex pairs: H simply sums them."""
    rng = np.random.default_rng(1)
    recs = []
    for _ in range(8):
        poses = synth.random_poses(rng, (5,), 6)
        edges = [se3_edge(rng, poses, 0, i) for i in range(1, 5)]
        edges += [se3_edge(rng, poses, 0, 2), se3_edge(rng, poses, 0, 2), se3_edge(rng, poses, 1, 3), se3_edge(rng, poses, 1, 3)]
        recs.append(R.pack_blanket(6, [7, 2, 4, 9, 11], poses, edges))
    for g, r in both(ctx, oracle, 6, recs, [4] * 8, R.ALG_NFR, R.TOPO_TREE):
        assert g["status"] == r["status"] == 0
        assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]]
        for eg, er in zip(g["edges"], r["edges"]):
            assert rel(eg["info"], er["info"]) <= 1e-9


@pytest.mark.parametrize("algorithm,topology", [(R.ALG_NFR, R.TOPO_TREE), (R.ALG_GLC, R.TOPO_DENSE), (R.ALG_GLC, R.TOPO_TREE)])
def test_several_removed_vertices(ctx, oracle, algorithm, topology):
    """Extended blankets (vertex_remover.cpp:93-95,142-195): m = d * #removed."""
    rng = np.random.default_rng(2)
    recs, nks = [], []
    for nrem in (2, 3):
        for _ in range(4):
            nk = 4
            n = nrem + nk
            poses = synth.random_poses(rng, (n,), 6)
            edges = [se3_edge(rng, poses, i, i + 1) for i in range(nrem - 1)]          # chain of removed vertices
            edges += [se3_edge(rng, poses, rng.integers(0, nrem), nrem + j) for j in range(nk)]
            edges += [se3_edge(rng, poses, nrem + j, nrem + (j + 1) % nk) for j in range(nk)]
            edges += [se3_edge(rng, poses, 0, nrem + 1)]
            recs.append(R.pack_blanket(6, list(range(100, 100 + n)), poses, edges, n_removed=nrem))
            nks.append(nk)
    for g, r in both(ctx, oracle, 6, recs, nks, algorithm, topology):
        assert g["status"] == r["status"] == 0
        assert g["n_edges"] == r["n_edges"]
        for eg, er in zip(g["edges"], r["edges"]):
            assert eg["v"] == er["v"]
            if algorithm == R.ALG_NFR:
                assert rel(eg["info"], er["info"]) <= 1e-9
            else:
                assert eg["rank"] == er["rank"]
                assert rel(eg["W"].T @ eg["W"], er["W"].T @ er["W"]) <= 1e-9


def test_weak_blanket_takes_choose_dimensions_branch(ctx, oracle):
    """More than d eigenvalues below the 1e-5 cutoff (an almost unconstrained kept vertex): the reference
    switches to chooseDimensions (logdet_function.cpp:42-60). The gauge shortcut must refuse the blanket
    and the eigen path must reproduce the oracle."""
    rng = np.random.default_rng(4)
    recs = []
    for _ in range(6):
        poses = synth.random_poses(rng, (5,), 6)
        edges = [se3_edge(rng, poses, 0, i) for i in range(1, 4)]
        edges.append(se3_edge(rng, poses, 0, 4, info=3e-6 * np.eye(6)))     # vertex 4 hangs on a 3e-6 thread
        edges.append(se3_edge(rng, poses, 1, 2))
        recs.append(R.pack_blanket(6, [9, 2, 4, 6, 8], poses, edges))
    for flags in (0, 1):
        for g, r in both(ctx, oracle, 6, recs, [4] * 6, R.ALG_NFR, R.TOPO_TREE, flags=flags):
            assert g["status"] == r["status"] == 0
            assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]]
            for eg, er in zip(g["edges"], r["edges"]):
                assert rel(eg["info"], er["info"]) <= 1e-9     # measured on B200: <= 1.4e-12


def test_not_positive_definite_marginal_is_reported(ctx, oracle):
    """A removed vertex without information: LLT(Lambda_mm) fails in the reference (garbage follows); here
    the blanket is flagged and produces no edges instead of asserting."""
    rng = np.random.default_rng(5)
    poses = synth.random_poses(rng, (4,), 6)
    edges = [se3_edge(rng, poses, 0, i, info=np.zeros((6, 6))) for i in range(1, 4)]
    edges.append(se3_edge(rng, poses, 1, 2))
    recs = [R.pack_blanket(6, [5, 1, 2, 3], poses, edges)]
    (g, r), = both(ctx, oracle, 6, recs, [3], R.ALG_NFR, R.TOPO_TREE)
    assert g["status"] == R.__dict__.get("ST_NOT_PD", 1) == r["status"]
    assert g["n_edges"] == 0


def test_se2_large_blanket(ctx, oracle):
    """intel-sized SE2 blankets (up to 14 vertices in the first round)."""
    blk = synth.make_blankets(14, 16, dim=3, variant="ring", seed=77)
    nk = R.n_kept_of(blk["records"], blk["rec_off"])
    recs = [blk["records"][blk["rec_off"][b]:blk["rec_off"][b + 1]] for b in range(16)]
    for alg, topo in ((R.ALG_NFR, R.TOPO_TREE), (R.ALG_GLC, R.TOPO_TREE)):
        for g, r in both(ctx, oracle, 3, recs, list(nk), alg, topo):
            assert g["status"] == r["status"] == 0
            assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]]


@pytest.mark.parametrize("dim,n", [(6, 24), (6, 40), (3, 48), (3, 100)])
@pytest.mark.parametrize("algorithm,topology", [(R.ALG_NFR, R.TOPO_TREE), (R.ALG_GLC, R.TOPO_TREE), (R.ALG_GLC, R.TOPO_DENSE)])
def test_blankets_beyond_shared_memory(ctx, oracle, dim, n, algorithm, topology):
    """Blankets whose working set exceeds the 227 KB of shared memory (SE3 > 19, SE2 > 39 vertices; the hubs
    of intel / sphere under Dense topologies) run the same kernel over a global-memory workspace."""
    nb = 3
    blk = synth.make_blankets(n, nb, dim=dim, variant="ring", seed=n)
    nk = R.n_kept_of(blk["records"], blk["rec_off"])
    recs = [blk["records"][blk["rec_off"][b]:blk["rec_off"][b + 1]] for b in range(nb)]
    for g, r in both(ctx, oracle, dim, recs, list(nk), algorithm, topology):
        assert g["status"] == r["status"] == 0
        assert g["n_edges"] == r["n_edges"]
        for eg, er in zip(g["edges"], r["edges"]):
            assert eg["v"] == er["v"]
            if algorithm == R.ALG_NFR:
                assert rel(eg["info"], er["info"]) <= 1e-9
            else:
                assert eg["rank"] == er["rank"]
                assert rel(eg["W"].T @ eg["W"], er["W"].T @ er["W"]) <= 1e-9


@pytest.mark.parametrize("n", [9, 10, 12, 13])
@pytest.mark.parametrize("topology", [R.TOPO_TREE, R.TOPO_SUBGRAPH])
def test_pose_only_promise_runs_the_lean_kernel(ctx, oracle, n, topology):
    """SPG_OPT_POSE_EDGES_ONLY (flags = 2): NFR rounds with 48 < 6 n <= 80 take the 256-thread / two-CTAs-per-SM
    variant with a small third buffer. Same results as the oracle, on the shortcut and on the forced eigen path."""
    nb = 6
    blk = synth.make_blankets(n, nb, dim=6, variant="ring", seed=100 + n)
    nk = R.n_kept_of(blk["records"], blk["rec_off"])
    recs = [blk["records"][blk["rec_off"][b]:blk["rec_off"][b + 1]] for b in range(nb)]
    for flags in (2, 3):
        if topology == R.TOPO_SUBGRAPH and flags == 3:
            continue
        for g, r in both(ctx, oracle, 6, recs, list(nk), R.ALG_NFR, topology, flags=flags):
            assert g["status"] == r["status"] == 0
            assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]]
            for eg, er in zip(g["edges"], r["edges"]):
                assert rel(eg["info"], er["info"]) <= 1e-9


def test_chunked_host_round_equals_single_chunk(oracle, monkeypatch):
    """spg_remove_round cuts a large round into chunks and overlaps H2D / kernels / D2H on three streams. With the
    chunk size forced down (SPG_CHUNK_BYTES) a mixed round of a few MB goes through ~30 chunks: the output must be
    bit-identical to the single-chunk run and match the oracle."""
    from sparsifyposegraph_b200 import capi
    parts = [synth.make_blankets(n, 300, dim=6, variant="ring", seed=7 * n) for n in (3, 5, 9, 12, 16, 4, 8)]
    recs = []
    for blk in parts:
        recs += [blk["records"][blk["rec_off"][b]:blk["rec_off"][b + 1]] for b in range(len(blk["rec_off"]) - 1)]
    rng = np.random.default_rng(0)
    recs = [recs[i] for i in rng.permutation(len(recs))]     # sizes interleaved: every chunk has several buckets
    records, rec_off = R.concat_records(recs)
    nk = R.n_kept_of(records, rec_off)
    out_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, nk)
    opts = capi.make_opts(R.TOPO_TREE, 1, flags=2)
    c1 = capi.Context(0)
    ref = c1.remove_round(6, R.ALG_NFR, opts, records, rec_off, out_off)[0].copy()
    c1.close()
    monkeypatch.setenv("SPG_CHUNK_BYTES", str(256 * 1024))
    c2 = capi.Context(0)
    got = c2.remove_round(6, R.ALG_NFR, opts, records, rec_off, out_off)[0].copy()
    c2.close()
    assert np.array_equal(ref, got)
    sample = list(range(0, len(recs), 97))
    o = oracle.remove_round(6, R.ALG_NFR, oracle.make_opts(R.TOPO_TREE, 1), records, rec_off, out_off, 0)[0]
    for b in sample:
        g = R.parse_out(got, out_off, b, 6, R.ALG_NFR, R.TOPO_TREE, nk[b])
        r = R.parse_out(o, out_off, b, 6, R.ALG_NFR, R.TOPO_TREE, nk[b])
        assert g["status"] == r["status"] == 0
        assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]]
        for eg, er in zip(g["edges"], r["edges"]):
            assert rel(eg["info"], er["info"]) <= 1e-9


@pytest.mark.parametrize("n", [18, 24, 30, 36])
@pytest.mark.parametrize("flags", [0, 2])
def test_se2_mid_sized_blankets_all_variants(ctx, oracle, n, flags):
    """SE2 blankets of 18-36 vertices (N = 54-108): the 512-thread dual-sweep kernel, the lean two-per-SM kernel
    (pose-only promise), the 256-thread kernel and (N > 96) the blocked-Cholesky route."""
    nb = 4
    blk = synth.make_blankets(n, nb, dim=3, variant="ring", seed=300 + n)
    nk = R.n_kept_of(blk["records"], blk["rec_off"])
    recs = [blk["records"][blk["rec_off"][b]:blk["rec_off"][b + 1]] for b in range(nb)]
    for alg, topo in ((R.ALG_NFR, R.TOPO_TREE), (R.ALG_GLC, R.TOPO_TREE)):
        if alg == R.ALG_GLC and flags:
            continue
        for g, r in both(ctx, oracle, 3, recs, list(nk), alg, topo, flags=flags):
            assert g["status"] == r["status"] == 0
            assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]]
            if alg == R.ALG_NFR:
                for eg, er in zip(g["edges"], r["edges"]):
                    assert rel(eg["info"], er["info"]) <= 1e-9


def test_wide_round_with_refusals_through_the_phase_barriers(ctx, oracle, monkeypatch):
    """fast_kernel's sub-warp-group variants run a wide round as CTAs of several warps with a CTA-wide barrier at every
    stage boundary (spg_fast.cuh, SPG_PHASE); every thread must execute the same number of barriers per blanket
    whatever the blanket does. A 2999-blanket round (ragged tail: not a multiple of the blankets per CTA) mixes regular
    5-vertex blankets with the three ways of leaving the kernel early — refused at the door (two removed vertices),
    refused in S0 (five parallel edges to one neighbour) and refused after both sweeps (weak blanket: the gauge
    shortcut's guard) — which blanket_kernel then re-runs. The result must equal the oracle and, bit for bit, the
    one-warp-per-CTA launch of the same binary (SPG_FAST_WPC=1: no barriers)."""
    from sparsifyposegraph_b200 import capi
    rng = np.random.default_rng(77)
    recs, nks = [], []
    for b in range(2999):
        if b % 11 == 5:     # two removed vertices, three kept
            poses = synth.random_poses(rng, (5,), 6)
            edges = [se3_edge(rng, poses, 0, 1)] + [se3_edge(rng, poses, j % 2, 2 + j) for j in range(3)]
            edges += [se3_edge(rng, poses, 2, 3), se3_edge(rng, poses, 3, 4)]
            recs.append(R.pack_blanket(6, list(range(100, 105)), poses, edges, n_removed=2))
            nks.append(3)
            continue
        poses = synth.random_poses(rng, (5,), 6)
        if b % 13 == 7:     # five parallel edges between the removed vertex and one neighbour
            edges = [se3_edge(rng, poses, 0, 1) for _ in range(5)] + [se3_edge(rng, poses, 0, i) for i in range(2, 5)]
            edges += [se3_edge(rng, poses, 1, 2), se3_edge(rng, poses, 3, 4)]
        elif b % 7 == 3:    # weak blanket (see test_weak_blanket_takes_choose_dimensions_branch)
            edges = [se3_edge(rng, poses, 0, i) for i in range(1, 4)]
            edges.append(se3_edge(rng, poses, 0, 4, info=3e-6 * np.eye(6)))
            edges.append(se3_edge(rng, poses, 1, 2))
        else:               # star + ring
            edges = [se3_edge(rng, poses, 0, i) for i in range(1, 5)]
            edges += [se3_edge(rng, poses, i, i % 4 + 1) for i in range(1, 5)]
        recs.append(R.pack_blanket(6, [9, 2, 4, 6, 8], poses, edges))
        nks.append(4)
    records, rec_off = R.concat_records(recs)
    out_off = R.out_offsets(6, R.ALG_NFR, R.TOPO_TREE, 1.0, nks)
    opts = capi.make_opts(R.TOPO_TREE, 1)
    phased = ctx.remove_round(6, R.ALG_NFR, opts, records, rec_off, out_off)[0].copy()
    retried = ctx.last_retry_count
    assert retried >= 2999 // 11 + 2999 // 13      # the refused blankets went through the device-side retry list
    monkeypatch.setenv("SPG_FAST_WPC", "1")
    onewarp = ctx.remove_round(6, R.ALG_NFR, opts, records, rec_off, out_off)[0].copy()
    monkeypatch.delenv("SPG_FAST_WPC")
    assert ctx.last_retry_count == retried
    assert np.array_equal(phased, onewarp)
    ro = oracle.remove_round(6, R.ALG_NFR, oracle.make_opts(R.TOPO_TREE, 1), records, rec_off, out_off, 0)[0]
    worst = 0.0
    for b, nk in enumerate(nks):
        g = R.parse_out(phased, out_off, b, 6, R.ALG_NFR, R.TOPO_TREE, nk)
        r = R.parse_out(ro, out_off, b, 6, R.ALG_NFR, R.TOPO_TREE, nk)
        assert g["status"] == r["status"] == 0, (b, g["status"], r["status"])
        assert [e["v"] for e in g["edges"]] == [e["v"] for e in r["edges"]], b
        for eg, er in zip(g["edges"], r["edges"]):
            worst = max(worst, rel(eg["info"], er["info"]))
    assert worst <= 1e-9, worst
