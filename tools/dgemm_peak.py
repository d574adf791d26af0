"""Second witness for the FP64 roofline denominator: cuBLAS DGEMM throughput through torch.matmul
(library code, measurement only — nothing in the product path calls it) next to the library's own DFMA probe."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsifyposegraph_b200 import capi  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[f"dgemm_{n}_tflops"] = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
    ctx = capi.Context(0)
    out["dfma_probe_tflops"] = ctx.fp64_peak_tflops()
    ctx.close()
    out["note"] = "cuBLAS DGEMM (torch.matmul, fp64) vs the register-resident DFMA loop of spg_fp64_peak_probe; B200, no clock lock"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
