import sys, time, numpy as np
from sparsifyposegraph_b200 import capi, synth, records as R
from oracle import pyoracle as O
def build(rows, cols):
    poses, edges, meas, info = synth.make_grid_graph(rows, cols, dim=6)
    g = capi.Graph(dim=6)
    for i in range(len(poses)): g.add_vertex(i, poses[i])
    for e in range(len(edges)): g.add_edge(int(edges[e,0]), int(edges[e,1]), meas[e], info[e])
    return g
rows, cols, mod = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = build(rows, cols)
last = g.max_vertex_id
which = capi.decimate_global(last, last, 10)
if mod:
    r, c = np.divmod(which, cols)
    colour = (r % mod) * mod + (c % mod)
    which = which[np.lexsort((which, colour))]
opts = capi.make_opts(R.TOPO_TREE, R.LIN_GLOBAL)
capi.rounds_begin(g, which, opts, R.ALG_NFR)
nr=0; tplan=0; tor=0; widths=[]; maxnv=0
while True:
    t0=time.time(); rd = capi.round_next(g); tplan += time.time()-t0
    if rd is None: break
    nr+=1; widths.append(rd["n"])
    hdr = rd["records"][rd["rec_off"][:-1]].view(np.int32).reshape(-1,2)[:,0]
    maxnv = max(maxnv, int(hdr.max()))
    t0=time.time()
    out = O.remove_round(rd["dim"], rd["algorithm"], O.make_opts(R.TOPO_TREE, R.LIN_GLOBAL), rd["records"], rd["rec_off"], rd["out_off"], 0)[0]
    tor += time.time()-t0
    t0=time.time(); capi.round_apply(g, out); tplan += time.time()-t0
print(rows, cols, "colour mod", mod, "removed", len(which), "rounds", nr, "plan+apply %.2fs"%tplan, "oracle %.1fs"%tor, "max blanket", maxnv, "widths", widths[:8], "...", widths[-4:])
