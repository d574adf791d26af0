import sys,time
sys.path.insert(0,'tests')
import datasets
from sparsifyposegraph_b200 import capi, records as R
ctx=capi.Context(0)
g=capi.Graph(datasets.path("sphere"))
last=g.max_vertex_id
which=capi.decimate_global(last,last,2)
t=time.time()
st=g.marginalize(ctx,which,capi.make_opts(R.TOPO_DENSE,R.LIN_GLOBAL),R.ALG_GLC)
print("sphere GLC dense", st, time.time()-t)
