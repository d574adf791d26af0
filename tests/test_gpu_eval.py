"""Whole-graph evaluator and optimiser on the GPU: spg_graph_kld against the CPU harness
(tests/kld_harness.py, a scipy restatement of graph_wrapper_g2o.cpp:531-548 + utils.cpp:70-97), spg_graph_chi2 against the
oracle's edge errors, spg_graph_optimize by its fixed point (zero gradient, recovered ground truth)."""
import numpy as np
import pytest

import datasets
import kld_harness as K
from sparsifyposegraph_b200 import records as R, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sparsifyposegraph_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name,alg,sparsity", [("sphere", R.ALG_NFR, 2), ("sphere", R.ALG_GLC, 2), ("intel", R.ALG_GLC, 2),
                                               ("manhattan", R.ALG_NFR, 3)])
def test_gpu_kld_equals_cpu_harness(ctx, oracle, name, alg, sparsity):
    """BASELINE.json configs[2]: KLD of the sparsified graph against the full-graph marginal, GPU evaluator vs harness."""
    from sparsifyposegraph_b200 import capi
    full, g = capi.Graph(datasets.path(name)), capi.Graph(datasets.path(name))
    which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, sparsity)
    st = g.marginalize(ctx, which, capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), alg)
    assert st["n_failed"] == 0
    kld_gpu, terms = full.kld(ctx, g)
    kld_cpu, _ = K.full_graph_kld(oracle, g.dim, K.poses_of(full), full.edges(), K.poses_of(g), g.edges())
    print(f"{name} alg {alg}: KLD gpu {kld_gpu:.9f} cpu {kld_cpu:.9f}; {terms['n_keep']} kept + {terms['n_marginalized']} marginalised dims, "
          f"{terms['device_ms']:.1f} ms on the device, {terms['flops'] / terms['device_ms'] / 1e9:.2f} TFLOP/s")
    assert 0 < kld_cpu < 1e4
    assert abs(kld_gpu - kld_cpu) <= 1e-6 * abs(kld_cpu)
    assert terms["mahalanobis"] < 1e-15     # both graphs sit at the estimates of the file


def test_gpu_kld_of_a_graph_against_itself_is_zero(ctx):
    from sparsifyposegraph_b200 import capi
    a, b = capi.Graph(datasets.path("intel")), capi.Graph(datasets.path("intel"))
    kld, terms = a.kld(ctx, b)
    assert abs(kld) < 1e-7 and terms["n_marginalized"] == 0


def test_gpu_kld_mahalanobis_term(ctx, oracle):
    """Moving a kept vertex of the full graph changes only d^T Lambda_x d (estimateDifference, :550-575) ... and the
    linearisation point of the full graph; compare with the harness on both counts."""
    from sparsifyposegraph_b200 import capi
    full, g = capi.Graph(datasets.path("intel")), capi.Graph(datasets.path("intel"))
    which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, 2)
    g.marginalize(ctx, which, capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), R.ALG_NFR)
    keep = int(g.vertex_ids()[10])
    p = full.vertex_pose(keep)
    p[:2] += 0.05
    full.set_vertex_pose(keep, p)
    kld_gpu, terms = full.kld(ctx, g)
    ids_s, Hs = K.graph_information(oracle, 3, K.poses_of(g), g.edges())
    d = np.zeros(3 * len(ids_s))
    d[3 * ids_s.index(keep):3 * ids_s.index(keep) + 2] = 0.05
    maha = float(d @ (Hs @ d))
    assert abs(terms["mahalanobis"] - maha) <= 1e-9 * maha
    kld_cpu, _ = K.full_graph_kld(oracle, 3, K.poses_of(full), full.edges(), K.poses_of(g), g.edges())
    assert abs(kld_gpu - (kld_cpu + 0.5 * maha)) <= 1e-6 * abs(kld_gpu)


@pytest.mark.parametrize("name", ["intel", "sphere"])
def test_gpu_chi2_equals_oracle_errors(ctx, oracle, name):
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path(name))
    poses = K.poses_of(g)
    ref = 0.0
    for e in g.edges():
        err = oracle.edge_error(g.dim, e["meas"], poses[int(e["v"][0])], poses[int(e["v"][1])])
        ref += float(err @ e["info"] @ err)
    got = g.chi2(ctx)
    assert abs(got - ref) <= 1e-9 * ref


@pytest.mark.parametrize("dim", [3, 6])
def test_optimizer_recovers_noise_free_grid(ctx, dim):
    """Noise-free measurements, perturbed estimates: Levenberg-Marquardt must return to chi2 = 0 and to the ground truth
    (vertex 0 fixed)."""
    from sparsifyposegraph_b200 import capi
    poses, edges, meas, info = synth.make_grid_graph(8, 9, dim=dim, sigma_t=0.0, sigma_r=0.0)
    g = synth.fill_graph(capi.Graph(dim=dim), poses, edges, meas, info)
    rng = np.random.default_rng(5)
    for i in range(1, len(poses)):
        p = poses[i].copy()
        p[:2] += rng.normal(0, 0.05, 2)
        g.set_vertex_pose(i, p)
    assert g.chi2(ctx) > 1.0
    st = g.optimize(ctx)
    assert st["chi2_final"] < 1e-12 * max(st["chi2_initial"], 1.0) and st["dimensions"] == dim * (len(poses) - 1)
    for i in (1, 17, len(poses) - 1):
        assert np.allclose(g.vertex_pose(i), poses[i], atol=1e-7)


@pytest.mark.parametrize("name", ["intel", "sphere"])
def test_optimizer_reaches_a_stationary_point_on_datasets(ctx, oracle, name):
    """GraphWrapperG2O::optimize on the reference's datasets: chi2 falls, and at the result the gradient J^T Omega e
    (from the oracle's Jacobians) vanishes relative to its size at the start."""
    from sparsifyposegraph_b200 import capi
    g = capi.Graph(datasets.path(name))

    def gradient_norm():
        poses = K.poses_of(g)
        acc = {}
        for e in g.edges():
            a, b = int(e["v"][0]), int(e["v"][1])
            err = oracle.edge_error(g.dim, e["meas"], poses[a], poses[b])
            Ji, Jj = oracle.edge_jacobians(g.dim, e["meas"], poses[a], poses[b])
            acc[a] = acc.get(a, 0) + Ji.T @ e["info"] @ err
            acc[b] = acc.get(b, 0) + Jj.T @ e["info"] @ err
        return float(np.sqrt(sum(float(v @ v) for k, v in acc.items() if k != 0)))
    g0 = gradient_norm()
    st = g.optimize(ctx)
    print(name, st)
    assert st["chi2_final"] < st["chi2_initial"]
    assert abs(g.chi2(ctx) - st["chi2_final"]) <= 1e-9 * st["chi2_final"]
    assert gradient_norm() <= 1e-5 * g0
