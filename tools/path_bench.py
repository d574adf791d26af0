#!/usr/bin/env python
"""Perf lines of the OTHER device paths on the C4 blanket sweep (bench.py times the NFR Chow-Liu-tree closed form):
GLC tree, GLC dense (R6) and NFR Subgraph (R9, interior-point Newton fit). Device-resident inputs, CUDA events on the
library stream, one JSON line per (path, n): ms, vertices/s, algorithmic GFLOP/s by SURVEY.md §8(d) (F_asm + F_schur +
F_cl + F_glcT for GLC tree; for NFR Subgraph the measured mean Newton iteration count is printed beside the time — the
closed-form flop formula does not apply), fraction of the measured DFMA peak.

    python tools/path_bench.py [--blankets 20000] [--iter-blankets 2000] > profiles/r2_paths.json"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blankets", type=int, default=20000)
    ap.add_argument("--iter-blankets", type=int, default=2000)
    ap.add_argument("--sizes", default="3,4,5,6,8,12,16")
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    import torch
    from sparsifyposegraph_b200 import capi, records as R, synth
    ctx = capi.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))
    peak = ctx.fp64_peak_tflops(3)
    sizes = [int(x) for x in args.sizes.split(",")]
    paths = [("glc", R.ALG_GLC, "tree", R.TOPO_TREE, args.blankets), ("glc", R.ALG_GLC, "dense", R.TOPO_DENSE, args.blankets),
             ("nfr", R.ALG_NFR, "subgraph", R.TOPO_SUBGRAPH, args.iter_blankets)]
    for aname, alg, tname, topo, B in paths:
        for n in sizes:
            blk = synth.make_blankets(n, B, dim=6, variant="ring", seed=synth.SEED + n)
            nk = np.full(B, n - 1, dtype=np.int64)
            out_off = R.out_offsets(6, alg, topo, 1.0, nk)
            opts = capi.make_opts(topo, R.LIN_GLOBAL, 1.0, flags=2)
            d_rec = torch.from_numpy(blk["records"].view(np.int64)).cuda()
            d_ro, d_oo = torch.from_numpy(blk["rec_off"]).cuda(), torch.from_numpy(out_off).cuda()
            d_out = torch.zeros(int(out_off[-1]), dtype=torch.int64, device="cuda")
            best = None
            for rep in range(args.repeat + 1):
                with torch.cuda.stream(stream):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    ctx.remove_round_device(6, alg, opts, B, d_rec.data_ptr(), d_ro.data_ptr(), d_oo.data_ptr(), d_out.data_ptr(),
                                            n, blk["E"])
                    b.record()
                ctx.sync()
                ms = a.elapsed_time(b)
                if rep and (best is None or ms < best):
                    best = ms
            hdr = d_out[torch.from_numpy(out_off[:-1]).cuda()].cpu().numpy().view(np.int32).reshape(-1, 2)
            hdr1 = d_out[torch.from_numpy(out_off[:-1] + 1).cuda()].cpu().numpy().view(np.int32).reshape(-1, 2)
            n_ok = int((hdr[:, 0] == 0).sum())
            line = {"path": f"{aname}-{tname}", "n": n, "E": blk["E"], "blankets": B, "ms": best, "vertices_per_s": B / (best * 1e-3),
                    "blankets_ok": n_ok, "fp64_peak_tflops": peak}
            if aname == "glc" and tname == "tree":
                alg_f = synth.algorithmic_bytes_flops(n, blk["E"], 6, "glc")
                line["algorithmic_gflops"] = alg_f["flops"] * B / (best * 1e-3) / 1e9
                line["frac_of_fp64_peak"] = line["algorithmic_gflops"] / 1e3 / peak
            if aname == "nfr":
                line["mean_newton_iterations"] = float(hdr1[:, 0].mean())
                line["ms_per_blanket_iteration"] = best / max(B * float(hdr1[:, 0].mean()), 1e-9)
            print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
