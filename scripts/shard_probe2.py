import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from kaminpar_b200 import lp
from kaminpar_b200.dist import CudaBackend, ShardedLP
from kaminpar_b200.graph import rmat
from oracle import bindings as B
g = B.oracle_rearrange(rmat(16, 16, 3))[0]
ctx = lp.create_default_context(); ctx.partition.setup(g, 8, 0.03)
h = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine)); h.set_graph(g)
b = CudaBackend(h, torch.device('cuda',0)); b.set_shard(0,1)
sizes = [b.subround_cap(sg) for sg in range(b.num_subrounds())]
print(sizes)
deg = g.degrees()
print('deg<8', (deg<8).sum(), '8..31', ((deg>=8)&(deg<32)).sum(), '32..255', ((deg>=32)&(deg<256)).sum(), '256..2047', ((deg>=256)&(deg<2048)).sum(), '>=2048', (deg>=2048).sum())
