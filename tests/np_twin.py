"""Independent numpy/LAPACK twin of the blanket pipeline, used ONLY to cross-check the C++ oracle
(SURVEY.md §8c item 3). Follows the same reference files, but with LAPACK eigh/cholesky/slogdet
instead of the oracle's hand-written dense kernels, and finite-difference Jacobians instead of the
analytic ones when ``fd=True``.
"""
import heapq

import numpy as np


def fd_jacobians(orc, dim, z, xi, xj, h=1e-6):
    """Central differences of the oracle's own error function through the vertex oplus."""
    Ji = np.zeros((dim, dim))
    Jj = np.zeros((dim, dim))
    for c in range(dim):
        dp = np.zeros(dim)
        dp[c] = h
        ep = orc.edge_error(dim, z, orc.oplus(dim, xi, dp), xj)
        em = orc.edge_error(dim, z, orc.oplus(dim, xi, -dp), xj)
        Ji[:, c] = wrap_err(dim, ep - em) / (2 * h)
        ep = orc.edge_error(dim, z, xi, orc.oplus(dim, xj, dp))
        em = orc.edge_error(dim, z, xi, orc.oplus(dim, xj, -dp))
        Jj[:, c] = wrap_err(dim, ep - em) / (2 * h)
    return Ji, Jj


def wrap_err(dim, e):
    if dim == 3:
        e = e.copy()
        e[2] = (e[2] + np.pi) % (2 * np.pi) - np.pi
    return e


def assemble(orc, dim, poses, edges, fd=False):
    """edges: list of (i, j, meas, info)."""
    n = len(poses)
    H = np.zeros((dim * n, dim * n))
    for (i, j, z, info) in edges:
        if fd:
            Ji, Jj = fd_jacobians(orc, dim, z, poses[i], poses[j])
        else:
            Ji, Jj = orc.edge_jacobians(dim, z, poses[i], poses[j])
        J = np.zeros((dim, dim * n))
        J[:, dim * i:dim * i + dim] = Ji
        J[:, dim * j:dim * j + dim] = Jj
        H += J.T @ info @ J
    return H


def schur(H, m):
    A, Bm, D = H[:m, :m], H[:m, m:], H[m:, m:]
    T = D - Bm.T @ np.linalg.solve(A, Bm)
    return 0.5 * (T + T.T)


def chow_liu_weights(T, n, d):
    C = np.linalg.inv(T + np.eye(T.shape[0]))
    w = {}
    for i in range(n - 1):
        for j in range(i + 1, n):
            idx = list(range(d * i, d * i + d)) + list(range(d * j, d * j + d))
            lx = np.linalg.slogdet(C[np.ix_(idx[:d], idx[:d])])[1]
            ly = np.linalg.slogdet(C[np.ix_(idx[d:], idx[d:])])[1]
            lxy = np.linalg.slogdet(C[np.ix_(idx, idx)])[1]
            w[(i, j)] = lx + ly - lxy
    return w


def kruskal_order(w, n):
    """Accepted edges (descending weight) then rejected ones; no tie handling (random data)."""
    order = sorted(w.items(), key=lambda kv: -kv[1])
    comp = list(range(n))

    def find(a):
        while comp[a] != a:
            a = comp[a]
        return a
    acc, rej = [], []
    for (i, j), _ in order:
        a, b = find(i), find(j)
        if a != b:
            comp[b] = a
            acc.append((i, j))
        else:
            rej.append((i, j))
    return acc + rej


def pattern(T, n, d, topology, chord_ratio=1.0):
    m = int((1 + chord_ratio) * (n - 1))
    full = m >= n * (n - 1) // 2
    if n == 2:
        return [(0, 1)]
    if topology == 3 or (topology == 1 and full):
        return [(i, j) for i in range(n - 1) for j in range(i + 1, n)]
    order = kruskal_order(chow_liu_weights(T, n, d), n)
    return order[:n - 1] if topology == 0 else order[:m]


def nfr_closed_form(orc, dim, T, kept_poses, pairs):
    d = dim
    w, V = np.linalg.eigh(T)
    S = 1.0 / w[d:]
    U = V[:, d:]
    Sigma = (U * S) @ U.T
    Xs = []
    for (a, b) in pairs:
        z = orc.compose(dim, orc.inverse(dim, kept_poses[a]), kept_poses[b])
        Ji, Jj = orc.edge_jacobians(dim, z, kept_poses[a], kept_poses[b])
        J = np.zeros((d, T.shape[0]))
        J[:, d * a:d * a + d] = Ji
        J[:, d * b:d * b + d] = Jj
        Xs.append(np.linalg.inv(J @ Sigma @ J.T))
    return Xs, (w, U, S)


def projected_kld(T, JXJ, d):
    """LogdetFunction::value at X: 0.5 [tr(S A) - logdet A - logdet S - r], A = U^T JXJ U."""
    w, V = np.linalg.eigh(T)
    S = 1.0 / w[d:]
    U = V[:, d:]
    A = U.T @ JXJ @ U
    return 0.5 * (np.sum(np.diag(A) * S) - np.linalg.slogdet(A)[1] - np.sum(np.log(S)) - len(S))


def glc_W(orc, dim, target, poses):
    """getEdge: W with W^T W = J^T-congruent target (returns W, meas)."""
    n = len(poses)
    r, J = orc.glc_reparam(dim, np.asarray(poses), np.zeros(dim * n))
    _, J = orc.glc_reparam(dim, np.asarray(poses), r)
    iJ = np.linalg.inv(J)
    M = iJ.T @ target @ iJ
    w, V = np.linalg.eigh(0.5 * (M + M.T))
    keep = w >= 1e-8
    W = (V[:, keep] * np.sqrt(w[keep])).T
    return W, r, J


def joint_marginal(T, idx):
    rest = [i for i in range(T.shape[0]) if i not in idx]
    if not rest:
        return T[np.ix_(idx, idx)]
    A = T[np.ix_(idx, idx)]
    Bm = T[np.ix_(idx, rest)]
    D = T[np.ix_(rest, rest)]
    return A - Bm @ np.linalg.solve(D, Bm.T)


def pinv_psd(a):
    w, V = np.linalg.eigh(a)
    tol = np.finfo(float).eps * a.shape[0] * np.abs(w).max()
    inv = np.where(w > tol, 1.0 / np.where(w > tol, w, 1.0), 0.0)
    return (V * inv) @ V.T
