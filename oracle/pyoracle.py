"""ctypes binding of the CPU oracle (oracle/_build/libspg_oracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY. Import this only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. PARITY UNPINNED (see oracle/capi.cpp header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libspg_oracle.so")


class SparsityOptions(C.Structure):
    _fields_ = [("topology", C.c_int32), ("lin_point", C.c_int32), ("chord_ratio", C.c_double),
                ("include_intra_clique", C.c_int32), ("reserved", C.c_int32)]


class RoundIn(C.Structure):
    _fields_ = [("dim", C.c_int32), ("algorithm", C.c_int32), ("opts", SparsityOptions),
                ("n_blankets", C.c_int32), ("reserved", C.c_int32),
                ("rec_off", C.c_void_p), ("records", C.c_void_p), ("out_off", C.c_void_p)]


class RoundOut(C.Structure):
    _fields_ = [("out", C.c_void_p), ("dbg_target", C.c_void_p), ("dbg_target_off", C.c_void_p),
                ("dbg_weights", C.c_void_p), ("dbg_weights_off", C.c_void_p)]


class EdgeDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nv", C.c_int32), ("rows", C.c_int32), ("uid_major", C.c_int32),
                ("uid_minor", C.c_int32)]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_version.restype = C.c_char_p
        L.orc_remove_round.restype = C.c_double
        L.orc_remove_round.argtypes = [C.POINTER(RoundIn), C.POINTER(RoundOut), C.c_int]
        L.orc_graph_load_g2o.restype = C.c_void_p
        L.orc_graph_load_g2o.argtypes = [C.c_char_p]
        L.orc_graph_create.restype = C.c_void_p
        L.orc_graph_create.argtypes = [C.c_int]
        L.orc_graph_destroy.argtypes = [C.c_void_p]
        L.orc_graph_add_vertex.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_graph_add_edge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for f in ("orc_graph_dim", "orc_graph_num_vertices", "orc_graph_num_edges", "orc_graph_max_vertex_id",
                  "orc_graph_log_size"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_graph_vertex_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_graph_vertex_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_graph_marginalize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SparsityOptions), C.c_int]
        L.orc_graph_log_entry.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_graph_log_detail.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_graph_edge_desc.argtypes = [C.c_void_p, C.c_int, C.POINTER(EdgeDesc)]
        L.orc_graph_edge_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("orc_decimate_global", "orc_decimate_online"):
            getattr(L, f).argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_decimate_cluster.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_compute_substitute_edge.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                  C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.orc_edge_error.argtypes = [C.c_int] + [C.c_void_p] * 4
        L.orc_edge_jacobians.argtypes = [C.c_int] + [C.c_void_p] * 5
        L.orc_oplus.argtypes = [C.c_int] + [C.c_void_p] * 3
        L.orc_compose.argtypes = [C.c_int] + [C.c_void_p] * 3
        L.orc_inverse.argtypes = [C.c_int] + [C.c_void_p] * 2
        L.orc_sym_eig.argtypes = [C.c_int] + [C.c_void_p] * 3
        L.orc_ldlt_sumlogd.restype = C.c_double
        L.orc_ldlt_sumlogd.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_glc_reparam.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_logdet_eval.restype = C.c_double
        L.orc_logdet_eval.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_opts(topology=0, lin_point=1, chord_ratio=1.0, include_intra_clique=True, flags=0):
    return SparsityOptions(int(topology), int(lin_point), float(chord_ratio), int(bool(include_intra_clique)), int(flags))


def remove_round(dim, algorithm, opts, records, rec_off, out_off, n_threads=1, want_target=None, want_weights=None,
                 weights_in=None):
    """Oracle twin of spg_remove_round. Returns (out uint64, seconds, dbg_target, dbg_weights).
    weights_in (with opts.flags & 4 and want_weights offsets): Chow-Liu weights are taken from it."""
    records = np.ascontiguousarray(records, dtype=np.uint64)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    out_off = np.ascontiguousarray(out_off, dtype=np.int64)
    nb = len(rec_off) - 1
    out = np.zeros(int(out_off[-1]), dtype=np.uint64)
    rin = RoundIn(dim, algorithm, opts, nb, 0, _p(rec_off).value, _p(records).value, _p(out_off).value)
    tgt = wts = None
    rout = RoundOut(_p(out).value, None, None, None, None)
    if want_target is not None:
        want_target = np.ascontiguousarray(want_target, dtype=np.int64)
        tgt = np.zeros(int(want_target[-1]), dtype=np.float64)
        rout.dbg_target = _p(tgt).value
        rout.dbg_target_off = _p(want_target).value
    if want_weights is not None:
        want_weights = np.ascontiguousarray(want_weights, dtype=np.int64)
        wts = np.zeros(int(want_weights[-1]), dtype=np.float64)
        if weights_in is not None:
            wts[:] = weights_in
        rout.dbg_weights = _p(wts).value
        rout.dbg_weights_off = _p(want_weights).value
    secs = lib().orc_remove_round(C.byref(rin), C.byref(rout), int(n_threads))
    return out, secs, tgt, wts


class Graph:
    """Oracle graph: sequential VertexRemover::remove on a g2o-free multigraph."""

    def __init__(self, path=None, dim=None):
        L = lib()
        if path is not None:
            self.h = L.orc_graph_load_g2o(os.fsencode(path))
            if not self.h:
                raise IOError(path)
        else:
            self.h = L.orc_graph_create(int(dim))
        self.dim = L.orc_graph_dim(self.h)
        self.P = 3 if self.dim == 3 else 7

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_graph_destroy(self.h)
            self.h = None

    def add_vertex(self, vid, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64)
        lib().orc_graph_add_vertex(self.h, int(vid), _p(p))

    def add_edge(self, a, b, meas, info):
        m = np.ascontiguousarray(meas, dtype=np.float64)
        i = np.ascontiguousarray(np.asarray(info, dtype=np.float64).T)
        lib().orc_graph_add_edge(self.h, int(a), int(b), _p(m), _p(i))

    @property
    def num_vertices(self):
        return lib().orc_graph_num_vertices(self.h)

    @property
    def num_edges(self):
        return lib().orc_graph_num_edges(self.h)

    @property
    def max_vertex_id(self):
        return lib().orc_graph_max_vertex_id(self.h)

    def vertex_ids(self):
        ids = np.zeros(self.num_vertices, dtype=np.int32)
        lib().orc_graph_vertex_ids(self.h, _p(ids))
        return ids

    def vertex_pose(self, vid):
        p = np.zeros(self.P)
        lib().orc_graph_vertex_pose(self.h, int(vid), _p(p))
        return p

    def marginalize(self, which, opts, algorithm):
        w = np.ascontiguousarray(which, dtype=np.int32)
        return lib().orc_graph_marginalize(self.h, _p(w), len(w), C.byref(opts), int(algorithm))

    def log(self):
        L = lib()
        out = []
        meta = np.zeros(6, dtype=np.int32)
        kld = C.c_double()
        for i in range(L.orc_graph_log_size(self.h)):
            L.orc_graph_log_entry(self.h, i, _p(meta), C.byref(kld))
            ids = np.zeros(meta[1], dtype=np.int32)
            pat = np.zeros((meta[5], 2), dtype=np.int32)
            L.orc_graph_log_detail(self.h, i, _p(ids), _p(pat))
            out.append({"root": int(meta[0]), "blanket": ids, "n_removed": int(meta[2]), "status": int(meta[3]),
                        "newton_iters": int(meta[4]), "kld": kld.value, "pattern": pat})
        return out

    def edges(self):
        """List of dicts in canonical order (same shape as sparsifyposegraph_b200.Graph.edges())."""
        L = lib()
        res = []
        d = EdgeDesc()
        for i in range(self.num_edges):
            L.orc_graph_edge_desc(self.h, i, C.byref(d))
            ids = np.zeros(d.nv, dtype=np.int32)
            if d.kind == 0:
                meas = np.zeros(self.P)
                info = np.zeros(self.dim * self.dim)
            elif d.kind == 1:
                meas = np.zeros(self.dim * d.nv)
                info = np.zeros(d.rows * self.dim * d.nv)
            else:
                meas = np.zeros((d.rows // self.dim) * self.P)
                info = np.zeros(d.rows * d.rows)
            L.orc_graph_edge_data(self.h, i, _p(ids), _p(meas), _p(info))
            if d.kind == 0:
                info = info.reshape(self.dim, self.dim).T
            elif d.kind == 1:
                info = info.reshape(d.rows, self.dim * d.nv)
            else:
                info = info.reshape(d.rows, d.rows).T
            res.append({"kind": d.kind, "v": ids, "rows": d.rows, "uid": (d.uid_major, d.uid_minor),
                        "meas": meas, "info": info})
        return res


def decimate_global(last, endvert, sparsity):
    out = np.zeros(max(endvert + 1, 1), dtype=np.int32)
    n = lib().orc_decimate_global(last, endvert, sparsity, _p(out), len(out))
    return out[:n]


def decimate_online(last, endvert, sparsity):
    out = np.zeros(4, dtype=np.int32)
    n = lib().orc_decimate_online(last, endvert, sparsity, _p(out), len(out))
    return out[:n]


def decimate_cluster(last, endvert, sparsity, cluster):
    out = np.zeros(max(endvert + 1, 1), dtype=np.int32)
    n = lib().orc_decimate_cluster(last, endvert, sparsity, cluster, _p(out), len(out))
    return out[:n]


def edge_error(dim, z, xi, xj):
    z, xi, xj = (np.ascontiguousarray(a, dtype=np.float64) for a in (z, xi, xj))
    e = np.zeros(dim)
    lib().orc_edge_error(dim, _p(z), _p(xi), _p(xj), _p(e))
    return e


def edge_jacobians(dim, z, xi, xj):
    z, xi, xj = (np.ascontiguousarray(a, dtype=np.float64) for a in (z, xi, xj))
    Ji = np.zeros(dim * dim)
    Jj = np.zeros(dim * dim)
    lib().orc_edge_jacobians(dim, _p(z), _p(xi), _p(xj), _p(Ji), _p(Jj))
    return Ji.reshape(dim, dim).T.copy(), Jj.reshape(dim, dim).T.copy()


def oplus(dim, x, delta):
    x, delta = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(delta, dtype=np.float64)
    out = np.zeros(3 if dim == 3 else 7)
    lib().orc_oplus(dim, _p(x), _p(delta), _p(out))
    return out


def compose(dim, a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(3 if dim == 3 else 7)
    lib().orc_compose(dim, _p(a), _p(b), _p(out))
    return out


def inverse(dim, a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.zeros(3 if dim == 3 else 7)
    lib().orc_inverse(dim, _p(a), _p(out))
    return out


def sym_eig(A):
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    Af = np.asfortranarray(A)
    w = np.zeros(n)
    V = np.zeros((n, n), order="F")
    rc = lib().orc_sym_eig(n, _p(Af), _p(w), _p(V))
    return w, V, rc


def ldlt_sumlogd(A):
    Af = np.asfortranarray(np.asarray(A, dtype=np.float64))
    pos = C.c_int()
    s = lib().orc_ldlt_sumlogd(Af.shape[0], _p(Af), C.byref(pos))
    return s, bool(pos.value)


def glc_reparam(dim, poses, meas):
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    n = poses.shape[0]
    meas = np.ascontiguousarray(meas, dtype=np.float64)
    r = np.zeros(dim * n)
    J = np.zeros((dim * n, dim * n), order="F")
    lib().orc_glc_reparam(dim, n, _p(poses), _p(meas), _p(r), _p(J))
    return r, np.array(J)


def logdet_eval(target, mapping, rho, x, want_hessian=True):
    """mapping: list of measurements, each a list of (J [rows, cols], offset)."""
    T = np.asfortranarray(np.asarray(target, dtype=np.float64))
    k = T.shape[0]
    rows = np.array([m[0][0].shape[0] for m in mapping], dtype=np.int32)
    nblk = np.array([len(m) for m in mapping], dtype=np.int32)
    cols = np.array([J.shape[1] for m in mapping for J, _ in m], dtype=np.int32)
    offs = np.array([o for m in mapping for _, o in m], dtype=np.int32)
    Jdata = np.concatenate([np.asarray(J, dtype=np.float64).T.reshape(-1) for m in mapping for J, _ in m])
    n = int(np.sum(rows.astype(np.int64) ** 2))
    x = np.ascontiguousarray(x, dtype=np.float64)
    g = np.zeros(n)
    H = np.zeros((n, n), order="F") if want_hessian else None
    cf = C.c_int()
    f = lib().orc_logdet_eval(k, _p(T), len(mapping), _p(rows), _p(nblk), _p(cols), _p(offs), _p(Jdata), float(rho),
                              _p(x), _p(g), _p(H), C.byref(cf))
    return f, g, (np.array(H) if H is not None else None), bool(cf.value)
