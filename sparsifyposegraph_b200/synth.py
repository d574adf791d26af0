"""Synthetic inputs for the benchmark / parity sweeps.

Pure data generation (numpy): random SE2/SE3 poses, noisy relative-pose measurements and
information matrices. No part of the removal path is computed here.
"""
from __future__ import annotations

import numpy as np

from . import records as R

SEED = 20261018


# ---- vectorised quaternion / SE3 helpers (x y z w order, as in g2o files) -----------------------

def qmul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    qv = np.concatenate([v, np.zeros(v.shape[:-1] + (1,))], axis=-1)
    return qmul(qmul(q, qv), qconj(q))[..., :3]


def qnormalize(q):
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return np.where(q[..., 3:4] < 0, -q, q)


def se3_compose(a, b):
    t = a[..., :3] + qrot(a[..., 3:], b[..., :3])
    q = qnormalize(qmul(a[..., 3:], b[..., 3:]))
    return np.concatenate([t, q], axis=-1)


def se3_inverse(a):
    qi = qconj(a[..., 3:])
    return np.concatenate([-qrot(qi, a[..., :3]), qi], axis=-1)


def se3_exp_small(rng, shape, sigma_t, sigma_r):
    t = rng.normal(0, sigma_t, shape + (3,))
    r = rng.normal(0, sigma_r, shape + (3,))
    ang = np.linalg.norm(r, axis=-1, keepdims=True)
    ax = r / np.maximum(ang, 1e-300)
    q = np.concatenate([ax * np.sin(ang / 2), np.cos(ang / 2)], axis=-1)
    return np.concatenate([t, q], axis=-1)


def se2_compose(a, b):
    c, s = np.cos(a[..., 2]), np.sin(a[..., 2])
    x = a[..., 0] + c * b[..., 0] - s * b[..., 1]
    y = a[..., 1] + s * b[..., 0] + c * b[..., 1]
    th = a[..., 2] + b[..., 2]
    th = (th + np.pi) % (2 * np.pi) - np.pi
    return np.stack([x, y, th], axis=-1)


def se2_inverse(a):
    c, s = np.cos(a[..., 2]), np.sin(a[..., 2])
    x = -(c * a[..., 0] + s * a[..., 1])
    y = -(-s * a[..., 0] + c * a[..., 1])
    return np.stack([x, y, -a[..., 2]], axis=-1)


def random_poses(rng, shape, dim):
    if dim == 3:
        xy = rng.uniform(-5, 5, shape + (2,))
        th = rng.uniform(-np.pi, np.pi, shape + (1,))
        return np.concatenate([xy, th], axis=-1)
    t = rng.uniform(-5, 5, shape + (3,))
    q = qnormalize(rng.normal(size=shape + (4,)))
    return np.concatenate([t, q], axis=-1)


def random_info(rng, shape, dim):
    """Dataset-like diagonal (sphere.g2o: 10,10,10,400,400,100; intel.g2o: 500,500,5000) plus a random
    PSD part, so the matrices are full and differently conditioned per edge."""
    d = dim
    diag = np.array([10, 10, 10, 400, 400, 100.0]) if d == 6 else np.array([500, 500, 5000.0])
    A = rng.normal(size=shape + (d, d))
    scale = np.sqrt(diag)
    M = np.einsum("...ij,...kj->...ik", A, A) * 0.25 / d
    M = M * scale[:, None] * scale[None, :]
    return M + np.diag(diag)


def blanket_topology(n, variant):
    """Local edge list of a synthetic blanket: vertex 0 is removed, 1..n-1 kept.
    'star': E = n-1 ; 'ring': star + cycle over the kept vertices (E = 2(n-1) for n-1 >= 3)."""
    nk = n - 1
    edges = [(0, i) for i in range(1, n)]
    if variant == "ring":
        if nk >= 3:
            edges += [(1 + i, 1 + (i + 1) % nk) for i in range(nk)]
        elif nk == 2:
            edges += [(1, 2)]
    return np.asarray(edges, dtype=np.int32)


def make_blankets(n, B, dim=6, variant="star", seed=None, sigma_t=0.05, sigma_r=0.02):
    """C4: B synthetic blankets with n vertices (1 removed). Returns a dict with the packed
    records plus the raw arrays."""
    rng = np.random.default_rng(SEED + n if seed is None else seed)
    P = R.pose_words(dim)
    poses = random_poses(rng, (B, n), dim)
    topo = blanket_topology(n, variant)
    E = len(topo)
    # ids: kept ascending, removed id strictly between two kept ids (random slot), edges go from
    # the lower to the higher id like the shipped datasets
    kept_ids = 10 + 2 * np.arange(n - 1)
    slot = rng.integers(0, n, size=B)
    rem_id = 9 + 2 * slot
    ids = np.concatenate([rem_id[:, None], np.broadcast_to(kept_ids, (B, n - 1))], axis=1).astype(np.int32)
    ev = np.broadcast_to(topo[None], (B, E, 2)).copy()
    a, b = ev[..., 0], ev[..., 1]
    ida = np.take_along_axis(ids, a, axis=1)
    idb = np.take_along_axis(ids, b, axis=1)
    swap = ida > idb
    a2 = np.where(swap, b, a)
    b2 = np.where(swap, a, b)
    ev = np.stack([a2, b2], axis=-1).astype(np.int32)
    Xi = np.take_along_axis(poses, ev[..., 0][..., None], axis=1)
    Xj = np.take_along_axis(poses, ev[..., 1][..., None], axis=1)
    if dim == 6:
        noise = se3_exp_small(rng, (B, E), sigma_t, sigma_r)
        meas = se3_compose(se3_compose(se3_inverse(Xi), Xj), noise)
    else:
        noise = np.concatenate([rng.normal(0, sigma_t, (B, E, 2)), rng.normal(0, sigma_r, (B, E, 1))], axis=-1)
        meas = se2_compose(se2_compose(se2_inverse(Xi), Xj), noise)
    info = random_info(rng, (B, E), dim)
    rec, rec_off = R.pack_uniform_pose_blankets(dim, ids, poses, ev, meas, info)
    return {"dim": dim, "n": n, "B": B, "E": E, "records": rec, "rec_off": rec_off, "ids": ids, "poses": poses,
            "edge_v": ev, "meas": meas, "info": info, "P": P}


def algorithmic_bytes_flops(n, E, dim=6, algorithm="nfr"):
    """Per-blanket algorithmic bytes / flops of the fused tree path, SURVEY.md §8(d) formulas
    (stated for SE3; d enters where the survey has it)."""
    d = dim
    k = d * (n - 1)
    b_in = 56 * n + 232 * E
    b_out = (n - 2) * (1256 if algorithm == "glc" else 352) if n >= 2 else 0
    if n == 2:
        b_out = 1256 if algorithm == "glc" else 352
    f_asm = E * (10 * d ** 3 + 300)
    f_schur = d ** 3 / 3 + 2 * d * d * k + 2 * d * k * k
    pairs = (n - 1) * (n - 2) / 2
    f_cl = 0 if n <= 3 else 7 * k ** 3 / 3 + pairs * ((2 * d) ** 3 / 3 + 2 * d ** 3 / 3)
    if algorithm == "glc":
        kk = k - 2 * d
        f_top = (n - 2) * (kk ** 3 / 3 + 4 * d * kk ** 2 + 8 * d * d * kk + 13 * d ** 3 + (47 / 3) * (2 * d) ** 3)
    else:
        f_top = 9 * k ** 3 + 2 * k * k * (k - d) + (n - 2) * (12 * d ** 3 + 7 * d ** 3 / 3)
    return {"bytes": float(b_in + b_out), "flops": float(f_asm + f_schur + f_cl + f_top)}


# ---- C5: synthetic SE3 grid graph ----------------------------------------------------------------

def make_grid_graph(rows, cols, dim=6, seed=SEED, sigma_t=0.05, sigma_r=0.02):
    """rows x cols grid (row-major ids). Odometry along the raster order plus vertical edges:
    degree <= 4 like sphere.g2o. Returns (poses[V,P], edges[E,2], meas[E,P], info[E,d,d])."""
    rng = np.random.default_rng(seed)
    V = rows * cols
    r, c = np.divmod(np.arange(V), cols)
    if dim == 6:
        t = np.stack([c * 1.0, r * 1.0, 0.1 * np.sin(0.3 * c + 0.2 * r)], axis=-1) + rng.normal(0, 0.02, (V, 3))
        q = qnormalize(np.concatenate([rng.normal(0, 0.1, (V, 3)), np.ones((V, 1))], axis=-1))
        poses = np.concatenate([t, q], axis=-1)
    else:
        poses = np.stack([c * 1.0, r * 1.0, rng.normal(0, 0.3, V)], axis=-1) + np.concatenate(
            [rng.normal(0, 0.02, (V, 2)), np.zeros((V, 1))], axis=-1)
    e_h = np.stack([np.arange(V - 1), np.arange(1, V)], axis=-1)
    vmask = r < rows - 1
    e_v = np.stack([np.arange(V)[vmask], np.arange(V)[vmask] + cols], axis=-1)
    edges = np.concatenate([e_h, e_v]).astype(np.int32)
    Xi, Xj = poses[edges[:, 0]], poses[edges[:, 1]]
    E = len(edges)
    if dim == 6:
        meas = se3_compose(se3_compose(se3_inverse(Xi), Xj), se3_exp_small(rng, (E,), sigma_t, sigma_r))
    else:
        noise = np.concatenate([rng.normal(0, sigma_t, (E, 2)), rng.normal(0, sigma_r, (E, 1))], axis=-1)
        meas = se2_compose(se2_compose(se2_inverse(Xi), Xj), noise)
    info = random_info(rng, (E,), dim)
    return poses, edges, meas, info


def grid_removal_order(rows, cols, sparsity=10, colour_mod=4, order=None, seed=SEED):
    """Removal list of BASELINE.json configs[4] (synthetic grid, 90 % removal at sparsity 10): the ids
    globalDecimate selects (decimation.cpp:36-49: i in [4, V-1] with i % sparsity != 0). The reference removes in
    list order, so the order is part of the input:
      "raster"  ascending ids: wavefront rounds a handful of blankets wide;
      "colour"  by colour class (r % m, c % m), then id: rounds of thousands, but hub blankets of 40-70 vertices
                at the end of the removal;
      "random"  seeded permutation: ~45 rounds whatever the grid size, blankets stay below 18 vertices.
    order=None keeps the historical meaning of colour_mod (0: raster, m: colour)."""
    V = rows * cols
    ids = np.arange(4, V, dtype=np.int32)
    ids = ids[ids % sparsity != 0]
    if order is None:
        order = "colour" if colour_mod else "raster"
    if order == "colour":
        r, c = np.divmod(ids, cols)
        colour = (r % colour_mod) * colour_mod + (c % colour_mod)
        ids = ids[np.lexsort((ids, colour))]
    elif order == "random":
        ids = np.random.default_rng(seed).permutation(ids)
    elif order != "raster":
        raise ValueError(order)
    return np.ascontiguousarray(ids, dtype=np.int32)


def fill_graph(graph, poses, edges, meas, info):
    """add_vertex / add_edge of a generated graph into a product or oracle Graph (same call surface)."""
    for i in range(len(poses)):
        graph.add_vertex(i, poses[i])
    for e in range(len(edges)):
        graph.add_edge(int(edges[e, 0]), int(edges[e, 1]), meas[e], info[e])
    return graph
