"""Full-graph KLD (tests/kld_harness.py: graph_wrapper_g2o.cpp:531-548 + utils.cpp:70-97).

CPU part: known answers of the harness itself on oracle-sparsified graphs. GPU part (-m gpu): BASELINE.json
configs[2] — sphere.g2o sparsified through the CUDA path and through the oracle's sequential loop give the same
KLD against the true full-graph marginal, to 1e-6 relative."""
import numpy as np
import pytest

import datasets
import kld_harness as K
from sparsifyposegraph_b200 import records as R
from sparsifyposegraph_b200 import synth


def chain(oracle, n, seed=3):
    rng = np.random.default_rng(seed)
    poses = synth.random_poses(rng, (n,), 6)
    g = oracle.Graph(dim=6)
    for i in range(n):
        g.add_vertex(i, poses[i])
    for i in range(n - 1):
        z = synth.se3_compose(synth.se3_compose(synth.se3_inverse(poses[i]), poses[i + 1]), synth.se3_exp_small(rng, (), 0.05, 0.02))
        g.add_edge(i, i + 1, z, synth.random_info(rng, (), 6))
    return g


def test_kld_of_a_graph_against_itself_is_zero(oracle):
    g = chain(oracle, 8)
    kld, _ = K.full_graph_kld(oracle, 6, K.poses_of(g), g.edges(), K.poses_of(g), g.edges())
    assert abs(kld) < 1e-9


def test_chain_marginalisation_is_exact(oracle):
    """Removing interior vertices of a chain leaves a chain: the tree approximation is the true marginal up to the
    linearisation of the new measurement (test_marginalize_se3.cpp's scenario) -> KLD ~ 0 for NFR and GLC."""
    for alg in (R.ALG_NFR, R.ALG_GLC):
        full, g = chain(oracle, 9), chain(oracle, 9)
        assert g.marginalize([4, 6], oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), alg) == 0
        kld, _ = K.full_graph_kld(oracle, 6, K.poses_of(full), full.edges(), K.poses_of(g), g.edges())
        assert abs(kld) < 1e-6, kld


def test_intel_tree_kld_nfr_equals_glc(oracle):
    """On a Chow-Liu tree the NFR closed form and the GLC conditionals both realise the KLD-optimal tree
    distribution of the blanket: the sparsified graphs carry the same information and the same (positive) KLD."""
    full = oracle.Graph(datasets.path("intel"))
    last = full.max_vertex_id
    which = oracle.decimate_global(last, last, 3)
    klds, marg = [], None
    for alg in (R.ALG_NFR, R.ALG_GLC):
        g = oracle.Graph(datasets.path("intel"))
        assert g.marginalize(which, oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), alg) == 0
        kld, marg = K.full_graph_kld(oracle, 3, K.poses_of(full), full.edges(), K.poses_of(g), g.edges(), marg)
        klds.append(kld)
    assert klds[0] > 1.0 and klds[0] < 1e4          # scripts/checkcomplete.py:63-64: "final KLD > 10000 is odd"
    assert abs(klds[0] - klds[1]) <= 1e-6 * klds[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name,alg,sparsity", [("sphere", R.ALG_NFR, 2), ("sphere", R.ALG_GLC, 2), ("manhattan", R.ALG_NFR, 2)])
def test_full_graph_kld_gpu_equals_oracle(oracle, name, alg, sparsity):
    from sparsifyposegraph_b200 import capi
    ctx = capi.Context(0)
    full = oracle.Graph(datasets.path(name))
    o = oracle.Graph(datasets.path(name))
    g = capi.Graph(datasets.path(name))
    last = g.max_vertex_id
    which = capi.decimate_global(last, last, sparsity)
    st = g.marginalize(ctx, which, capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), alg)
    assert st["n_failed"] == 0
    assert o.marginalize(which, oracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), alg) == 0
    fp, fe = K.poses_of(full), full.edges()
    kld_o, marg = K.full_graph_kld(oracle, g.dim, fp, fe, K.poses_of(o), o.edges())
    kld_g, _ = K.full_graph_kld(oracle, g.dim, fp, fe, K.poses_of(g), g.edges(), marg)
    print(f"{name} alg {alg} sparsity {sparsity}: full-graph KLD oracle {kld_o:.9f} gpu {kld_g:.9f}")
    assert 0 < kld_o < 1e4
    assert abs(kld_g - kld_o) <= 1e-6 * abs(kld_o)
    ctx.close()
