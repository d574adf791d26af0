"""CPU-only look at the wavefront rounds of the synthetic grid workload (BASELINE.json configs[4]) under different
removal orders: number of rounds, round widths, largest blanket, scheduler time. The blankets are computed by the
ORACLE here, which is why this script lives under tests/ — it is for choosing a removal order, not a benchmark.
usage: python tests/grid_plan.py ROWS COLS ORDER   with ORDER = raster | colourM (M = modulus) | random | level"""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from sparsifyposegraph_b200 import capi, synth, records as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

rows, cols, order = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
g = synth.fill_graph(capi.Graph(dim=6), *synth.make_grid_graph(rows, cols, dim=6))
which = synth.grid_removal_order(rows, cols, 10, 0)
r, c = np.divmod(which, cols)
if order.startswith("colour"):
    m = int(order[6:])
    which = which[np.lexsort((which, (r % m) * m + (c % m)))]
elif order == "random":
    which = np.random.default_rng(1).permutation(which).astype(np.int32)
elif order == "level":
    # multigrid-like: vertices on the coarsest lattices go last (level = trailing zeros of r and c)
    def tz(x):
        x = np.where(x == 0, 1 << 20, x)
        return np.log2(x & -x).astype(int)
    lvl = np.minimum(tz(r), tz(c))
    which = which[np.lexsort((which, (r % 4) * 4 + (c % 4), lvl))]
which = np.ascontiguousarray(which, dtype=np.int32)
opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
capi.rounds_begin(g, which, opts, R.ALG_NFR)
nr = 0
tplan = tor = 0.0
widths, sizes = [], []
while True:
    t0 = time.time()
    rd = capi.round_next(g)
    tplan += time.time() - t0
    if rd is None:
        break
    nr += 1
    widths.append(rd["n"])
    sizes.append(rd["records"][rd["rec_off"][:-1]].view(np.int32).reshape(-1, 2)[:, 0])
    t0 = time.time()
    out = O.remove_round(rd["dim"], rd["algorithm"], O.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rd["records"], rd["rec_off"], rd["out_off"], 0)[0]
    tor += time.time() - t0
    t0 = time.time()
    capi.round_apply(g, out)
    tplan += time.time() - t0
sizes = np.concatenate(sizes)
print(f"{rows}x{cols} {order}: {len(which)} removals, {nr} rounds (widest {max(widths)}), plan+apply {tplan:.2f} s, oracle {tor:.1f} s, "
      f"largest blanket {sizes.max()} vertices, {int((sizes > 19).sum())} blankets beyond 19 vertices, {int((sizes > 16).sum())} beyond 16")
