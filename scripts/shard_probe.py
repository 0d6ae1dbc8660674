import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from kaminpar_b200 import lp
from kaminpar_b200.dist import CudaBackend, ShardedLP
from kaminpar_b200.graph import rmat
from oracle import bindings as B
g = B.oracle_rearrange(rmat(int(sys.argv[1]), 16, 3))[0]
print('n', g.n, 'm', g.m, 'maxdeg', g.degrees().max())
ctx = lp.create_default_context(); ctx.partition.setup(g, 8, 0.03)
mcw = lp.compute_max_cluster_weight(ctx.coarsening, ctx.partition, g.n, g.total_node_weight())
dev = torch.device('cuda', 0)
for rank, world in ((0,1),(1,2),(0,2)):
    h = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine)); h.set_graph(g)
    drv = ShardedLP(CudaBackend(h, dev), g.n, 5, rank, world)
    # emulate: run only the sweep of each sub-round for this rank's shard, commit with own buffer replicated
    drv.world = 1  # no collectives; commit uses [send] only -> exercises slicing kernels for rank/world
    drv.b.lib.kmp_lp_set_shard(h._h, rank, world)
    b = drv.b
    b.begin_cluster(mcw, None); b.begin_iteration()
    for sg in range(b.num_subrounds()):
        cap, size = b.subround_cap(sg)
        if size == 0: continue
        send = b.alloc(4 + 2*cap); recv = b.alloc((4+2*cap)*world); recv.zero_()
        b.sweep(0, sg, send); recv[:4+2*cap] = send
        b.commit(0, sg, recv)
    print(rank, world, 'moved', b.end_iteration()); torch.cuda.synchronize()
print('done')
