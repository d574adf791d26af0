"""Full-graph KLD harness (test infrastructure, CPU, numpy/scipy) — the quality metric of BASELINE.json
configs[2] ("sphere ... KLD vs full-graph marginal").

Restates GraphWrapperG2O::kullbackLeibler (reference src/graph_wrapper_g2o.cpp:531-548, computeIndices
:472-499) and kullbackLeiblerDivergence (src/utils.cpp:70-97, InformationInformation mode):

    Lambda_y = marginal of the FULL graph's information onto the variables the sparsified graph keeps
               (sparse Cholesky of the marginalised block, Schur complement),
    Lambda_x = information of the sparsified graph,
    KLD      = 1/2 [ tr(Lambda_y^-1 Lambda_x) + d^T Lambda_x d - logdet Lambda_x + logdet Lambda_y - n ].

Both informations are H = sum_e J^T Omega J at the graphs' current estimates with vertex 0 fixed (g2o's
_Hpp, :382-396); no optimiser runs here, so both graphs are linearised at the estimates of the file and the
mean difference d is zero unless the caller moved vertices. Edge Jacobians come from the oracle's
restatement of g2o (oracle/poses.hpp, FD-checked in tests/test_oracle.py); GLC factors contribute
J = W * J_reparam (src/glc_edge.cpp:40-49) with Omega = I.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _edge_blocks(oracle, dim, e, poses):
    """[(vertex id, J block (rows x dim))], Omega (rows x rows) of one edge dict (capi/pyoracle Graph.edges())."""
    v = [int(x) for x in e["v"]]
    if e["kind"] == 0:
        Ji, Jj = oracle.edge_jacobians(dim, e["meas"], poses[v[0]], poses[v[1]])
        return [(v[0], Ji), (v[1], Jj)], np.asarray(e["info"])
    if e["kind"] == 1:
        _, Jr = oracle.glc_reparam(dim, np.stack([poses[i] for i in v]), e["meas"])
        J = np.asarray(e["info"]) @ Jr                      # W (rank x d nv) * J_reparam
        return [(vi, J[:, dim * q:dim * (q + 1)]) for q, vi in enumerate(v)], np.eye(J.shape[0])
    # MultiEdgeCorrelated (multi_edge_correlated.hpp:96-140): stacked pose Jacobians, full information
    P = 3 if dim == 3 else 7
    nm = e["rows"] // dim
    meas = np.asarray(e["meas"]).reshape(nm, P)
    pairs = np.asarray(e["pairs"]).reshape(nm, 2)
    blocks = {vi: np.zeros((e["rows"], dim)) for vi in v}
    for m in range(nm):
        a, b = v[pairs[m, 0]], v[pairs[m, 1]]
        Ji, Jj = oracle.edge_jacobians(dim, meas[m], poses[a], poses[b])
        blocks[a][dim * m:dim * (m + 1)] += Ji
        blocks[b][dim * m:dim * (m + 1)] += Jj
    return [(vi, blocks[vi]) for vi in v], np.asarray(e["info"])


def graph_information(oracle, dim, poses, edges, fixed=0):
    """Sparse information matrix of a graph: poses {id: flat pose}, edges as returned by Graph.edges().
    Returns (ids in matrix order, csc matrix). Vertex `fixed` is excluded (g2o: vertex 0 setFixed)."""
    ids = sorted(i for i in poses if i != fixed)
    pos = {vid: dim * q for q, vid in enumerate(ids)}
    rows, cols, vals = [], [], []
    for e in edges:
        blocks, Om = _edge_blocks(oracle, dim, e, poses)
        for va, Ja in blocks:
            if va == fixed:
                continue
            M = Ja.T @ Om
            for vb, Jb in blocks:
                if vb == fixed:
                    continue
                B = M @ Jb
                r, c = np.meshgrid(np.arange(dim) + pos[va], np.arange(dim) + pos[vb], indexing="ij")
                rows.append(r.ravel())
                cols.append(c.ravel())
                vals.append(B.ravel())
    n = dim * len(ids)
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsc()
    return ids, H


def true_marginal(ids_full, H_full, keep_ids, dim, chunk=512):
    """Dense Schur complement of the full-graph information onto the kept vertices (graph_wrapper_g2o.cpp:539-542)."""
    keep = set(keep_ids)
    ik = np.concatenate([np.arange(dim) + dim * q for q, v in enumerate(ids_full) if v in keep])
    im = np.concatenate([np.arange(dim) + dim * q for q, v in enumerate(ids_full) if v not in keep] or [np.zeros(0, int)])
    ik, im = ik.astype(int), im.astype(int)
    Hkk = H_full[ik][:, ik].toarray()
    if len(im) == 0:
        return Hkk
    Hmm = H_full[im][:, im].tocsc()
    Hmk = H_full[im][:, ik].tocsc()
    lu = spla.splu(Hmm)
    for c0 in range(0, len(ik), chunk):
        rhs = Hmk[:, c0:c0 + chunk].toarray()
        Hkk[:, c0:c0 + chunk] -= Hmk.T @ lu.solve(rhs)
    return 0.5 * (Hkk + Hkk.T)


def kld_information_information(diff, info_x, info_y):
    """kullbackLeiblerDivergence(diff, infox, maty, InformationInformation), src/utils.cpp:70-97 (dense)."""
    cx = sla.cho_factor(info_x, lower=True)
    cy = sla.cho_factor(info_y, lower=True)
    logdetx = 2.0 * np.sum(np.log(np.diag(cx[0])))
    logdety = -2.0 * np.sum(np.log(np.diag(cy[0])))
    innerprod = np.trace(sla.cho_solve(cy, info_x))
    maha = float(diff @ info_x @ diff) if diff is not None else 0.0
    return 0.5 * (innerprod + maha - logdetx - logdety - info_x.shape[0])


def full_graph_kld(oracle, dim, full_poses, full_edges, sparse_poses, sparse_edges, marginal=None):
    """KLD(true marginal of the full graph || sparsified graph). `marginal` caches (ids, Lambda_y)."""
    ids_s, Hs = graph_information(oracle, dim, sparse_poses, sparse_edges)
    if marginal is None:
        ids_f, Hf = graph_information(oracle, dim, full_poses, full_edges)
        marginal = (ids_s, true_marginal(ids_f, Hf, ids_s, dim))
    assert marginal[0] == ids_s
    return kld_information_information(None, Hs.toarray(), marginal[1]), marginal


def poses_of(graph):
    return {int(i): graph.vertex_pose(int(i)) for i in graph.vertex_ids()}
