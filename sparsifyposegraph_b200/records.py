"""Packed blanket records (numpy side).

Mirrors ``include/spg_capi.h`` / ``include/spg_record.h``: one wavefront round is a flat uint64
buffer of per-blanket input records plus word offsets, and a flat uint64 output buffer. These
helpers only build / parse buffers; all arithmetic of the removal path happens in the CUDA
library (``capi.py``) — there is no numpy implementation of the path in this package.
"""
from __future__ import annotations

import numpy as np

EDGE_POSE, EDGE_GLC, EDGE_MULTI = 0, 1, 2
ALG_NFR, ALG_GLC = 0, 1
TOPO_TREE, TOPO_SUBGRAPH, TOPO_CLIQUEY_SUBGRAPH, TOPO_DENSE, TOPO_CLIQUEY_DENSE = range(5)
LIN_LOCAL, LIN_GLOBAL = 0, 1
REC_HEADER_WORDS = 4
OUT_HEADER_WORDS = 4

STATUS_NAMES = {
    0: "OK", 1: "NOT_PD_MARGINAL", 2: "NOT_PD_CHOWLIU", 3: "EIG_NOCONV", 4: "NOT_PD_CLOSED",
    5: "TOO_LARGE", 6: "LINESEARCH_FAIL", 7: "KLD_INF", 8: "UNSUPPORTED", 9: "NOT_PD_JOINT",
}


def pad2(n):
    return (n + 1) // 2


def pose_words(dim: int) -> int:
    return 3 if dim == 3 else 7


def edge_words(dim, kind, nv, rows):
    w = 2 + pad2(nv)
    if kind == EDGE_POSE:
        w += pose_words(dim) + dim * dim
    elif kind == EDGE_GLC:
        w += dim * nv + rows * dim * nv
    else:
        nmeas = rows // dim
        w += pad2(2 * nmeas) + nmeas * pose_words(dim) + rows * rows
    return w


def out_edge_count(algorithm, topology, chord_ratio, n_kept):
    """spgr_out_edge_count (pseudo_chow_liu.cpp:41-86, topology_provider_glc.cpp:113-183)."""
    n = int(n_kept)
    if algorithm == ALG_GLC:
        if n <= 0:
            return 0
        if n == 1 or topology == TOPO_DENSE:
            return 1
        return n
    if n < 2:
        return 0
    if n == 2:
        return 1
    m = int((1 + chord_ratio) * (n - 1))
    allp = n * (n - 1) // 2
    if topology == TOPO_TREE:
        return n - 1
    if topology == TOPO_SUBGRAPH:
        return allp if m >= allp else m
    if topology == TOPO_DENSE:
        return allp
    if topology == TOPO_CLIQUEY_DENSE:
        return 1
    return n - 1


def out_slot_words(dim, algorithm, topology, n_kept):
    P = pose_words(dim)
    if algorithm == ALG_GLC:
        nvcap = n_kept if (topology == TOPO_DENSE or n_kept == 1) else 2
        c = dim * nvcap
        return 1 + pad2(nvcap) + c + c * c
    if topology in (TOPO_CLIQUEY_DENSE, TOPO_CLIQUEY_SUBGRAPH):
        return 0   # variable-size entries (spg_record.h: spgr_out_entry_words)
    return 1 + P + dim * dim


def out_record_words(dim, algorithm, topology, chord_ratio, n_kept):
    if algorithm == ALG_NFR and topology in (TOPO_CLIQUEY_DENSE, TOPO_CLIQUEY_SUBGRAPH):
        nm = max(int(n_kept) - 1, 0)
        w = OUT_HEADER_WORDS + nm * (5 + pose_words(dim)) + (nm * dim) ** 2
        return (w + 1) & ~1
    w = OUT_HEADER_WORDS + out_edge_count(algorithm, topology, chord_ratio, n_kept) * out_slot_words(
        dim, algorithm, topology, n_kept)
    return (w + 1) & ~1


def out_offsets(dim, algorithm, topology, chord_ratio, n_kept):
    """Word offsets [B+1] of the output records for blankets with n_kept[b] kept vertices."""
    n_kept = np.asarray(n_kept, dtype=np.int64)
    uniq = np.unique(n_kept)
    words = np.zeros(n_kept.shape, dtype=np.int64)
    for u in uniq:
        words[n_kept == u] = out_record_words(dim, algorithm, topology, chord_ratio, int(u))
    off = np.zeros(len(n_kept) + 1, dtype=np.int64)
    np.cumsum(words, out=off[1:])
    return off


# ------------------------------------------------------------------------------------------------
# packing
# ------------------------------------------------------------------------------------------------

def pack_uniform_pose_blankets(dim, ids, poses, edge_v, meas, info, n_removed=1, tags=None):
    """Vectorised packer for B blankets that share (n_vert, n_edges) and hold only POSE edges.

    ids    [B, n] int     original vertex ids, removed first then kept ascending
    poses  [B, n, P]      SE2 (x y theta) / SE3 (t, qx qy qz qw)
    edge_v [E, 2] or [B, E, 2] local vertex indices (from, to)
    meas   [B, E, P]
    info   [B, E, d, d]   row/col symmetric information
    returns (records uint64 [B*W], rec_off int64 [B+1])
    """
    ids = np.asarray(ids)
    B, n = ids.shape
    P, d = pose_words(dim), dim
    edge_v = np.asarray(edge_v)
    if edge_v.ndim == 2:
        edge_v = np.broadcast_to(edge_v[None], (B,) + edge_v.shape)
    E = edge_v.shape[1]
    poses_off = REC_HEADER_WORDS + pad2(n)
    etab_off = poses_off + n * P
    e0 = etab_off + pad2(E)
    We = 2 + 1 + P + d * d
    W = e0 + E * We
    W += W & 1
    rec = np.zeros((B, W), dtype=np.uint64)
    i32 = rec.view(np.int32)
    f64 = rec.view(np.float64)
    i32[:, 0] = n
    i32[:, 1] = n_removed
    i32[:, 2] = E
    i32[:, 3] = dim
    i32[:, 4] = W
    i32[:, 6] = np.arange(B) if tags is None else tags
    i32[:, 2 * REC_HEADER_WORDS:2 * REC_HEADER_WORDS + n] = ids
    f64[:, poses_off:poses_off + n * P] = np.asarray(poses, dtype=np.float64).reshape(B, n * P)
    eoffs = e0 + We * np.arange(E)
    i32[:, 2 * etab_off:2 * etab_off + E] = eoffs
    meas = np.asarray(meas, dtype=np.float64)
    info = np.asarray(info, dtype=np.float64)
    for e in range(E):
        o = int(eoffs[e])
        i32[:, 2 * o] = EDGE_POSE
        i32[:, 2 * o + 1] = 2
        i32[:, 2 * o + 2] = d
        i32[:, 2 * (o + 2)] = edge_v[:, e, 0]
        i32[:, 2 * (o + 2) + 1] = edge_v[:, e, 1]
        f64[:, o + 3:o + 3 + P] = meas[:, e]
        f64[:, o + 3 + P:o + 3 + P + d * d] = info[:, e].transpose(0, 2, 1).reshape(B, d * d)
    rec_off = np.arange(B + 1, dtype=np.int64) * W
    return rec.reshape(-1), rec_off


def pack_blanket(dim, ids, poses, edges, n_removed=1, tag=0):
    """General (slow, python) packer of ONE blanket. ``edges`` is a list of dicts:
       POSE : {"kind": 0, "v": [i, j], "meas": [P], "info": [d, d]}
       GLC  : {"kind": 1, "v": [...], "meas": [d*nv], "W": [rows, d*nv]}
       MULTI: {"kind": 2, "v": [...], "pairs": [[a, b], ...], "meas": [nmeas, P], "info": [rows, rows]}
    """
    P, d = pose_words(dim), dim
    n, E = len(ids), len(edges)
    kinds = [int(e.get("kind", 0)) for e in edges]
    nvs = [len(e["v"]) for e in edges]
    rows = []
    for e, k in zip(edges, kinds):
        if k == EDGE_POSE:
            rows.append(d)
        elif k == EDGE_GLC:
            rows.append(int(np.asarray(e["W"]).shape[0]))
        else:
            rows.append(d * len(e["pairs"]))
    poses_off = REC_HEADER_WORDS + pad2(n)
    etab_off = poses_off + n * P
    e0 = etab_off + pad2(E)
    ew = [edge_words(dim, k, nv, r) for k, nv, r in zip(kinds, nvs, rows)]
    W = e0 + sum(ew)
    W += W & 1
    rec = np.zeros(W, dtype=np.uint64)
    i32 = rec.view(np.int32)
    f64 = rec.view(np.float64)
    i32[0:7] = [n, n_removed, E, dim, W, 0, tag]
    i32[2 * REC_HEADER_WORDS:2 * REC_HEADER_WORDS + n] = ids
    f64[poses_off:poses_off + n * P] = np.asarray(poses, dtype=np.float64).reshape(-1)
    o = e0
    for ei, e in enumerate(edges):
        i32[2 * etab_off + ei] = o
        k, nv, r = kinds[ei], nvs[ei], rows[ei]
        i32[2 * o] = k
        i32[2 * o + 1] = nv
        i32[2 * o + 2] = r
        i32[2 * (o + 2):2 * (o + 2) + nv] = e["v"]
        p = o + 2 + pad2(nv)
        if k == EDGE_POSE:
            f64[p:p + P] = e["meas"]
            f64[p + P:p + P + d * d] = np.asarray(e["info"], dtype=np.float64).T.reshape(-1)
        elif k == EDGE_GLC:
            c = d * nv
            f64[p:p + c] = e["meas"]
            f64[p + c:p + c + r * c] = np.asarray(e["W"], dtype=np.float64).reshape(-1)
        else:
            nm = len(e["pairs"])
            i32[2 * p:2 * p + 2 * nm] = np.asarray(e["pairs"], dtype=np.int32).reshape(-1)
            p2 = p + pad2(2 * nm)
            f64[p2:p2 + nm * P] = np.asarray(e["meas"], dtype=np.float64).reshape(-1)
            f64[p2 + nm * P:p2 + nm * P + r * r] = np.asarray(e["info"], dtype=np.float64).T.reshape(-1)
        o += ew[ei]
    return rec


def concat_records(recs):
    off = np.zeros(len(recs) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in recs], out=off[1:])
    return (np.concatenate(recs) if recs else np.zeros(0, np.uint64)), off


def record_header(records, rec_off, b):
    h = records[rec_off[b]:rec_off[b] + 2].view(np.int32)
    return {"n_vert": int(h[0]), "n_removed": int(h[1]), "n_edges": int(h[2]), "dim": int(h[3])}


def n_kept_of(records, rec_off):
    """n_kept per blanket, read from the record headers."""
    i32 = records.view(np.int32)
    base = 2 * np.asarray(rec_off[:-1])
    return (i32[base] - i32[base + 1]).astype(np.int64)


# ------------------------------------------------------------------------------------------------
# parsing outputs
# ------------------------------------------------------------------------------------------------

def parse_out(out, out_off, b, dim, algorithm, topology, n_kept):
    """Decode output record b into a dict."""
    w = out[out_off[b]:out_off[b + 1]]
    i32 = w.view(np.int32)
    f64 = w.view(np.float64)
    res = {"status": int(i32[0]), "n_edges": int(i32[1]), "newton_iters": int(i32[2]), "flags": int(i32[3]),
           "kld": float(f64[2]), "edges": []}
    P, d = pose_words(dim), dim
    slot = out_slot_words(dim, algorithm, topology, n_kept)
    o = OUT_HEADER_WORDS
    if algorithm == ALG_NFR and topology in (TOPO_CLIQUEY_DENSE, TOPO_CLIQUEY_SUBGRAPH):
        # correlated topologies: a sequence of variable-size entries (nmeas == 1: pose edge, > 1: multi-edge)
        for _ in range(res["n_edges"]):
            nm, rows = int(i32[2 * o]), int(i32[2 * o + 1])
            ab = [int(x) for x in i32[2 * (o + 1):2 * (o + 1) + 2 * nm]]
            m0 = o + 1 + pad2(2 * nm)
            meas = f64[m0:m0 + nm * P].reshape(nm, P).copy()
            info = f64[m0 + nm * P:m0 + nm * P + rows * rows].reshape(rows, rows).T.copy()
            res["edges"].append({"v": ab, "pairs": [(ab[2 * q], ab[2 * q + 1]) for q in range(nm)], "nmeas": nm, "meas": meas,
                                 "info": info})
            o = m0 + nm * P + rows * rows
        return res
    for _ in range(res["n_edges"]):
        if algorithm == ALG_NFR:
            a, bb = int(i32[2 * o]), int(i32[2 * o + 1])
            meas = f64[o + 1:o + 1 + P].copy()
            info = f64[o + 1 + P:o + 1 + P + d * d].reshape(d, d).T.copy()
            res["edges"].append({"v": [a, bb], "meas": meas, "info": info})
        else:
            nvcap = n_kept if (topology == TOPO_DENSE or n_kept == 1) else 2
            c = d * nvcap
            nv, rank = int(i32[2 * o]), int(i32[2 * o + 1])
            v = [int(x) for x in i32[2 * (o + 1):2 * (o + 1) + nv]]
            m0 = o + 1 + pad2(nvcap)
            meas = f64[m0:m0 + d * nv].copy()
            Wfull = f64[m0 + c:m0 + c + c * c].reshape(c, c)
            res["edges"].append({"v": v, "rank": rank, "meas": meas, "W": Wfull[:rank, :d * nv].copy()})
        o += slot
    return res


def nfr_uniform_view(out, out_off, dim, n_kept, n_edges):
    """Vectorised view of NFR outputs of B uniform blankets -> (status[B], pairs[B,E,2], meas[B,E,P], info[B,E,d,d])."""
    B = len(out_off) - 1
    W = int(out_off[1] - out_off[0])
    o = out[:B * W].reshape(B, W)
    i32 = o.view(np.int32)
    f64 = o.view(np.float64)
    P, d = pose_words(dim), dim
    slot = 1 + P + d * d
    status = i32[:, 0].copy()
    pairs = np.zeros((B, n_edges, 2), dtype=np.int32)
    meas = np.zeros((B, n_edges, P))
    info = np.zeros((B, n_edges, d, d))
    for e in range(n_edges):
        s = OUT_HEADER_WORDS + e * slot
        pairs[:, e, 0] = i32[:, 2 * s]
        pairs[:, e, 1] = i32[:, 2 * s + 1]
        meas[:, e] = f64[:, s + 1:s + 1 + P]
        info[:, e] = f64[:, s + 1 + P:s + 1 + P + d * d].reshape(B, d, d).transpose(0, 2, 1)
    return status, pairs, meas, info
