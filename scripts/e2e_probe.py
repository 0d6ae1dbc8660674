"""Where does an e2e step go? set_graph (H2D + work lists) vs cluster (compute + D2H), with / without the upload overlap."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kaminpar_b200 import lp
from kaminpar_b200.graph import CSRGraph

dev = torch.device("cuda", 0)
xadj64, adj64, k = bench.generate("rmat22", dev)
n, m = xadj64.numel() - 1, adj64.numel()
h_xadj = torch.empty(n + 1, dtype=torch.int32, pin_memory=True).copy_(xadj64.to(torch.int32))
h_adj = torch.empty(m, dtype=torch.int32, pin_memory=True).copy_(adj64.to(torch.int32))
h_out = torch.empty(n, dtype=torch.int32, pin_memory=True)
g = CSRGraph.__new__(CSRGraph)
g.xadj = h_xadj.numpy().view(np.uint32); g.adjncy = h_adj.numpy().view(np.uint32); g.vwgt = None; g.adjwgt = None; g.sorted = True; g.buckets = None
ctx = lp.create_default_context(); ctx.partition.setup(g, k, 0.03)
mcw = lp.compute_max_cluster_weight(ctx.coarsening, ctx.partition, n, n)
h = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine))
out = h_out.numpy().view(np.uint32)
t0 = time.perf_counter(); d = torch.empty(m, dtype=torch.int32, device=dev); d.copy_(h_adj, non_blocking=True); torch.cuda.synchronize(); print("torch H2D adjncy ms", (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); d.copy_(h_adj, non_blocking=True); torch.cuda.synchronize(); print("torch H2D adjncy ms (2nd)", (time.perf_counter() - t0) * 1e3)
for rep in range(4):
    t0 = time.perf_counter(); h.set_graph(g); torch.cuda.synchronize(); t1 = time.perf_counter()
    _, st = h.cluster(mcw, out=out); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: set_graph {1e3*(t1-t0):.2f} ms, cluster {1e3*(t2-t1):.2f} ms (device {st.device_ms:.2f})")
