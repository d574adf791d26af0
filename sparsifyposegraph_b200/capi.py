"""ctypes binding of libspg_b200.so (include/spg_capi.h).

The CUDA library is the product; this module only marshals numpy buffers / raw device pointers
into the C ABI. There is no CPU implementation behind it: if the shared library is missing or no
sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspg_b200.so")


class SparsityOptions(C.Structure):
    """reference src/sparsity_options.h:11-30"""
    _fields_ = [("topology", C.c_int32), ("lin_point", C.c_int32), ("chord_ratio", C.c_double),
                ("include_intra_clique", C.c_int32), ("reserved", C.c_int32)]


class RoundIn(C.Structure):
    _fields_ = [("dim", C.c_int32), ("algorithm", C.c_int32), ("opts", SparsityOptions),
                ("n_blankets", C.c_int32), ("reserved", C.c_int32),
                ("rec_off", C.c_void_p), ("records", C.c_void_p), ("out_off", C.c_void_p)]


class RoundOut(C.Structure):
    _fields_ = [("out", C.c_void_p), ("dbg_target", C.c_void_p), ("dbg_target_off", C.c_void_p),
                ("dbg_weights", C.c_void_p), ("dbg_weights_off", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("reserved", C.c_int32), ("max_record_words", C.c_int64),
                ("max_out_words", C.c_int64)]


class EdgeDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nv", C.c_int32), ("rows", C.c_int32), ("uid_major", C.c_int32),
                ("uid_minor", C.c_int32)]


class MarginalizeStats(C.Structure):
    _fields_ = [("n_rounds", C.c_int32), ("n_blankets", C.c_int32), ("max_round_width", C.c_int32),
                ("max_blanket_vertices", C.c_int32), ("n_failed", C.c_int32), ("n_dropped_edges", C.c_int32),
                ("pack_ms", C.c_double), ("gpu_ms", C.c_double), ("splice_ms", C.c_double)]


class SpgError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpgError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the node-removal path)")
        L = C.CDLL(LIB_PATH)
        L.spg_version.restype = C.c_char_p
        L.spg_last_error.restype = C.c_char_p
        L.spg_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.spg_destroy.argtypes = [C.c_void_p]
        L.spg_launch_count.restype = C.c_int64
        L.spg_launch_count.argtypes = [C.c_void_p]
        L.spg_last_kernel_ms.restype = C.c_double
        L.spg_last_kernel_ms.argtypes = [C.c_void_p]
        L.spg_remove_round.argtypes = [C.c_void_p, C.POINTER(RoundIn), C.POINTER(RoundOut)]
        L.spg_remove_round_device.argtypes = [C.c_void_p, C.POINTER(RoundIn), C.POINTER(RoundOut), C.c_int32, C.c_int32]
        L.spg_sync.argtypes = [C.c_void_p]
        L.spg_stream.restype = C.c_void_p
        L.spg_stream.argtypes = [C.c_void_p]
        L.spg_out_record_words.restype = C.c_int64
        L.spg_out_record_words.argtypes = [C.c_int32, C.c_int32, C.POINTER(SparsityOptions), C.c_int32]
        L.spg_fp64_peak_probe.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def make_opts(topology=0, lin_point=1, chord_ratio=1.0, include_intra_clique=True, flags=0):
    return SparsityOptions(int(topology), int(lin_point), float(chord_ratio), int(bool(include_intra_clique)), int(flags))


def _check(rc):
    if rc != 0:
        raise SpgError(f"spg status {rc}: {lib().spg_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p).value if a is not None else None


class Context:
    """spg_ctx: one CUDA device, one stream, growable device staging buffers."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        cfg = Config(int(device), 0, 0, 0)
        _check(lib().spg_create(C.byref(self.h), C.byref(cfg)))

    def close(self):
        if getattr(self, "h", None):
            lib().spg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def launches(self):
        return int(lib().spg_launch_count(self.h))

    @property
    def last_kernel_ms(self):
        return float(lib().spg_last_kernel_ms(self.h))

    def stream(self):
        return lib().spg_stream(self.h)

    def remove_round(self, dim, algorithm, opts, records, rec_off, out_off, out=None, want_target=None,
                     want_weights=None):
        """Host buffers in / out (H2D + kernels + D2H inside). Returns (out, dbg_target, dbg_weights)."""
        records = np.ascontiguousarray(records, dtype=np.uint64)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
        out_off = np.ascontiguousarray(out_off, dtype=np.int64)
        nb = len(rec_off) - 1
        if out is None:
            out = np.empty(int(out_off[-1]), dtype=np.uint64)
        rin = RoundIn(dim, algorithm, opts, nb, 0, _p(rec_off), _p(records), _p(out_off))
        rout = RoundOut(_p(out), None, None, None, None)
        tgt = wts = None
        if want_target is not None:
            want_target = np.ascontiguousarray(want_target, dtype=np.int64)
            tgt = np.zeros(int(want_target[-1]), dtype=np.float64)
            rout.dbg_target, rout.dbg_target_off = _p(tgt), _p(want_target)
        if want_weights is not None:
            want_weights = np.ascontiguousarray(want_weights, dtype=np.int64)
            wts = np.zeros(int(want_weights[-1]), dtype=np.float64)
            rout.dbg_weights, rout.dbg_weights_off = _p(wts), _p(want_weights)
        _check(lib().spg_remove_round(self.h, C.byref(rin), C.byref(rout)))
        return out, tgt, wts

    def remove_round_device(self, dim, algorithm, opts, n_blankets, d_records, d_rec_off, d_out_off, d_out,
                            max_n_vert, max_n_edges):
        """Device pointers (ints) already resident in HBM; asynchronous on the context stream."""
        rin = RoundIn(dim, algorithm, opts, int(n_blankets), 0, int(d_rec_off), int(d_records), int(d_out_off))
        rout = RoundOut(int(d_out), None, None, None, None)
        _check(lib().spg_remove_round_device(self.h, C.byref(rin), C.byref(rout), int(max_n_vert), int(max_n_edges)))

    def sync(self):
        _check(lib().spg_sync(self.h))

    def fp64_peak_tflops(self, repeats=5):
        """Measured DFMA peak of this device (roofline denominator for the FP64-bound kernels)."""
        v = C.c_double()
        _check(lib().spg_fp64_peak_probe(self.h, int(repeats), C.byref(v)))
        return float(v.value)
