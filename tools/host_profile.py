"""Host-side profile of the round scheduler WITHOUT a GPU (development tool): planning + packing
(spg_graph_round_next) and splicing (spg_graph_round_apply) are the product's code, the blankets in between are
computed by the CPU oracle standing in for the device (checker code, never shipped on this path). Prints the seconds
spent in each of the two host calls; SPG_HOST_PROF=1 adds the scheduler's own breakdown on stderr.

    SPG_HOST_PROF=1 python tools/host_profile.py --rows 400 --cols 400
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402
from sparsifyposegraph_b200 import capi, records as R, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=400)
    ap.add_argument("--cols", type=int, default=400)
    ap.add_argument("--sparsity", type=int, default=10)
    ap.add_argument("--order", default="random", choices=["random", "colour", "raster"])
    ap.add_argument("--oracle-threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    pyoracle.build()
    data = synth.make_grid_graph(args.rows, args.cols, dim=6)
    g = synth.fill_graph(capi.Graph(dim=6), *data)
    which = synth.grid_removal_order(args.rows, args.cols, args.sparsity, 4, args.order)
    opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    oopts = pyoracle.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
    L = capi.lib()
    capi._round_protos(L)
    t0 = time.perf_counter()
    capi.rounds_begin(g, which, opts, R.ALG_NFR)
    t_begin = time.perf_counter() - t0
    t_next = t_apply = t_dev = 0.0
    rounds = widest = 0
    while True:
        r = capi.RoundIn()
        t0 = time.perf_counter()
        capi._check(L.spg_graph_round_next(g.h, C.byref(r)))
        t_next += time.perf_counter() - t0
        n = r.n_blankets
        if n == 0:
            break
        rounds += 1
        widest = max(widest, n)
        rec_off = np.ctypeslib.as_array(C.cast(r.rec_off, C.POINTER(C.c_int64)), shape=(n + 1,))
        out_off = np.ctypeslib.as_array(C.cast(r.out_off, C.POINTER(C.c_int64)), shape=(n + 1,))
        records = np.ctypeslib.as_array(C.cast(r.records, C.POINTER(C.c_uint64)), shape=(int(rec_off[-1]),))
        t0 = time.perf_counter()
        out = pyoracle.remove_round(r.dim, r.algorithm, oopts, records, rec_off, out_off, args.oracle_threads)[0]
        t_dev += time.perf_counter() - t0
        t0 = time.perf_counter()
        capi._check(L.spg_graph_round_apply(g.h, capi._p(out)))
        t_apply += time.perf_counter() - t0
    print(json.dumps({"grid": [args.rows, args.cols], "removed": int(len(which)), "rounds": rounds, "widest": widest,
                      "begin_s": t_begin, "plan_pack_s": t_next, "splice_s": t_apply, "host_s": t_next + t_apply,
                      "host_vertices_per_s": len(which) / (t_next + t_apply), "oracle_standin_s": t_dev,
                      "host_threads": os.environ.get("SPG_HOST_THREADS", "default"), "cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
