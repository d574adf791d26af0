#!/usr/bin/env python
"""Throughput on the reference's own datasets (BASELINE.json configs[0..2]), graph level:
spg_graph_marginalize (plan + pack + H2D + kernels + D2H + splice, every wavefront round one spg_remove_round) with the
CPU restatement of the reference path (oracle/, sequential one-at-a-time loop like VertexRemover::remove) timed beside it
on the same removal list.

    python tools/dataset_bench.py [--repeat 5] > profiles/r2_datasets.json

One JSON line per job: vertices/s on the GPU path (best and median of --repeat fresh graphs), rounds, widest round, split
of the time (host planning/packing, GPU incl. copies, splicing), the oracle's vertices/s on one host thread, and the
worst relative Frobenius difference of the resulting information matrices (the parity number of tests/test_gpu_graph.py).
These workloads are latency-bound: sphere has 25 blankets per round, the tails of intel / manhattan a handful — the
per-round cost (a kernel launch per size bucket + two copies + a stream sync) dominates, not the FP64 pipe."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

JOBS = [  # (config, dataset, algorithm, topology, sparsity)
    ("C1", "intel", "glc", "tree", 2),
    ("C1", "intel", "glc", "dense", 2),
    ("C2", "manhattan", "nfr", "tree", 2),
    ("C2", "manhattan", "nfr", "subgraph", 2),
    ("C3", "sphere", "glc", "tree", 2),
    ("C3", "sphere", "nfr", "tree", 2),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated dataset:algorithm:topology filters, e.g. intel:glc:tree")
    args = ap.parse_args()
    import datasets
    from test_gpu_graph import compare_graphs
    from sparsifyposegraph_b200 import capi, records as R
    from oracle import pyoracle
    pyoracle.build()
    ctx = capi.Context(0)
    ALG = {"glc": R.ALG_GLC, "nfr": R.ALG_NFR}
    TOPO = {"tree": R.TOPO_TREE, "dense": R.TOPO_DENSE, "subgraph": R.TOPO_SUBGRAPH}
    only = [f for f in args.only.split(",") if f]
    for cfg, name, alg, topo, sparsity in JOBS:
        if only and f"{name}:{alg}:{topo}" not in only:
            continue
        path = datasets.path(name)
        opts = capi.make_opts(TOPO[topo], R.LIN_GLOBAL)
        times, st, g = [], None, None
        for rep in range(args.repeat + 1):  # first pass = warm-up (module load, buffer growth)
            g = capi.Graph(path)
            which = capi.decimate_global(g.max_vertex_id, g.max_vertex_id, sparsity)
            launches0 = ctx.launches
            t0 = time.perf_counter()
            st = g.marginalize(ctx, which, opts, ALG[alg])
            dt = time.perf_counter() - t0
            launches = ctx.launches - launches0
            if rep:
                times.append(dt)
        line = {"config": cfg, "dataset": name, "algorithm": alg, "topology": topo, "sparsity": sparsity,
                "removed": int(len(which)), "vertices": int(g.num_vertices + len(which)), "rounds": st["n_rounds"],
                "max_round_width": st["max_round_width"], "max_blanket_vertices": st["max_blanket_vertices"],
                "kernel_launches": int(launches),
                "gpu_path": {"vertices_per_s_best": len(which) / min(times), "vertices_per_s_median": len(which) / float(np.median(times)),
                             "ms_best": 1e3 * min(times), "ms_per_round": 1e3 * min(times) / max(st["n_rounds"], 1),
                             "split_ms_last": {"plan_pack": st["pack_ms"], "gpu_incl_copies": st["gpu_ms"], "splice": st["splice_ms"]}},
                "lin_point": "global: the estimates stored in the file (no optimiser on this path)"}
        if not args.no_oracle:
            o = pyoracle.Graph(path)
            t0 = time.perf_counter()
            bad = o.marginalize(which, pyoracle.make_opts(TOPO[topo], R.LIN_GLOBAL), ALG[alg])
            dt = time.perf_counter() - t0
            line["cpu_oracle"] = {"vertices_per_s": len(which) / dt, "ms": 1e3 * dt, "threads": 1, "failed": int(bad),
                                  "kind": "port (restatement of the reference's sequential loop; the reference itself cannot be built offline)"}
            line["speedup_vs_cpu_oracle"] = dt / min(times)
            try:
                line["worst_rel_frobenius_vs_oracle"] = float(compare_graphs(g, o, tol=1e-6))
            except AssertionError as e:  # reported, not hidden
                line["parity_error"] = str(e)[:200]
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
