"""N > 1 path on CPU: the sharded driver (kaminpar_b200/dist.py) under the `gloo` backend with
world_size 2 and 3, using an oracle-backed stand-in for the per-rank sweep / commit so that no GPU
is needed. Checks the exchange protocol (packed proposal buffers, all_gather, replicated commit,
favored fix-up): the result on every rank must equal the single-process oracle `sync` result."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kaminpar_b200.dist import ShardedLP  # noqa: E402
from kaminpar_b200.graph import random_weights, rmat  # noqa: E402
from oracle import bindings as B  # noqa: E402


class OracleBackend:
    """CPU stand-in for CudaBackend: same stepping interface, decisions computed by the oracle
    (lpo_sync_select_all / lpo_sync_commit)."""

    def __init__(self, g, seed, S=8, G=4, passes=1):
        self.g, self.seed, self.S, self.G, self.passes = g, seed, S, G, passes
        self.lib = B.oracle()
        sg = np.zeros(g.n, np.uint32)
        self.lib.lpo_sync_subround_index(C.c_uint32(g.n), g.xadj.ctypes.data_as(C.c_void_p), C.c_int(seed),
                                         C.c_uint32(S), C.c_uint32(G), C.c_uint32(0xFFFFFFFF),
                                         sg.ctypes.data_as(C.c_void_p))
        self.sg_of = sg
        self.call = 0

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world

    def num_subrounds(self):
        return 4 * self.S

    def _members(self, sg):
        return np.nonzero(self.sg_of == sg)[0].astype(np.uint32)

    def subround_cap(self, sg):
        size = int((self.sg_of == sg).sum())
        return (size + self.world - 1) // self.world + 1, size

    def alloc(self, words):
        return torch.zeros(words, dtype=torch.int32)

    def begin_cluster(self, mcw, communities):
        g = self.g
        self.mode, self.mcw = 0, mcw
        self.labels = np.arange(g.n, dtype=np.uint32)
        self.weights = (np.ones(g.n, np.int32) if g.vwgt is None else g.vwgt.copy())
        self.favored = np.arange(g.n, dtype=np.uint32)
        self.active = np.ones(g.n, np.uint8)
        self.max_w = None
        self.min_w = None

    def begin_refine(self, k, max_bw, min_bw, communities, partition):
        g = self.g
        self.mode, self.mcw = 1, 0
        self.labels = np.ascontiguousarray(partition, np.uint32).copy()
        w = np.ones(g.n, np.int64) if g.vwgt is None else g.vwgt.astype(np.int64)
        self.weights = np.bincount(self.labels, weights=w, minlength=k).astype(np.int32)
        self.active = np.ones(g.n, np.uint8)
        self.max_w = np.ascontiguousarray(max_bw, np.int32)
        self.min_w = None if min_bw is None else np.ascontiguousarray(min_bw, np.int32)

    def begin_iteration(self):
        self.moved = 0

    def sweep(self, it, sg, send):
        mem = self._members(sg)
        mine = mem[self.rank::self.world]  # any disjoint cover of the list gives the same result
        mine = mine[self.active[mine] != 0]
        tgt, fav = B.oracle_sync_select_all(self.mode, self.g, self.labels, self.weights, max_weights=self.max_w,
                                            max_cluster_weight=self.mcw, min_weights=self.min_w, seed=self.seed,
                                            call=self.call, iteration=it)
        self.active[mine] = 0
        moving = mine[tgt[mine] != self.labels[mine]]
        staying = mine[tgt[mine] == self.labels[mine]]
        if self.mode == 0:
            ok = fav[staying] != 0xFFFFFFFF
            self.favored[staying[ok]] = fav[staying[ok]]
        cap = (send.numel() - 4) // 2
        buf = send.numpy()
        buf[0] = len(moving)
        buf[4:4 + len(moving)] = moving.view(np.int32)
        buf[4 + cap:4 + cap + len(moving)] = tgt[moving].view(np.int32)

    def commit(self, it, sg, gathered):
        buf = gathered.numpy()
        words = buf.size // self.world
        cap = (words - 4) // 2
        us, ts = [], []
        for r in range(self.world):
            cnt = int(buf[r * words])
            us.append(buf[r * words + 4: r * words + 4 + cnt].view(np.uint32))
            ts.append(buf[r * words + 4 + cap: r * words + 4 + cap + cnt].view(np.uint32))
        pu = np.ascontiguousarray(np.concatenate(us), np.uint32)
        pt = np.ascontiguousarray(np.concatenate(ts), np.uint32)
        g = self.g
        self.lib.lpo_sync_commit.restype = C.c_uint32
        self.moved += self.lib.lpo_sync_commit(
            C.c_uint32(g.n), g.xadj.ctypes.data_as(C.c_void_p), g.adjncy.ctypes.data_as(C.c_void_p),
            B._opt(g.vwgt, np.int32), C.c_uint32(len(pu)), pu.ctypes.data_as(C.c_void_p),
            pt.ctypes.data_as(C.c_void_p), self.labels.ctypes.data_as(C.c_void_p),
            self.weights.ctypes.data_as(C.c_void_p), C.c_uint32(len(self.weights)), B._opt(self.max_w, np.int32),
            C.c_int32(self.mcw), B._opt(self.min_w, np.int32), self.active.ctypes.data_as(C.c_void_p),
            C.c_int(self.seed), C.c_uint32(self.call), C.c_uint32(it), C.c_uint32(sg), C.c_uint32(self.passes))

    def end_iteration(self):
        return self.moved

    def favored_export(self, buf):
        buf.numpy()[:] = (self.favored ^ np.arange(self.g.n, dtype=np.uint32)).view(np.int32)

    def favored_import(self, buf):
        self.favored = buf.numpy().view(np.uint32) ^ np.arange(self.g.n, dtype=np.uint32)

    def finish(self, n, k=None, fetch=True):
        if self.mode == 0:
            p = B.oracle_params(B.default_cluster_params(), self.S, self.G, self.passes)
            g = self.g
            self.lib.lpo_sync_post_passes(
                C.c_uint32(g.n), g.xadj.ctypes.data_as(C.c_void_p), g.adjncy.ctypes.data_as(C.c_void_p),
                B._opt(g.vwgt, np.int32), self.labels.ctypes.data_as(C.c_void_p),
                self.weights.ctypes.data_as(C.c_void_p), self.favored.ctypes.data_as(C.c_void_p), C.c_int32(self.mcw),
                C.byref(p))
            self.call += 1
        return self.labels.copy(), (self.weights.copy() if self.mode == 1 else None), None


def _worker(rank, world, port, graph_seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = B.oracle_rearrange(random_weights(rmat(11, 8, graph_seed), 3, max_vwgt=3))[0]
        mcw = B.oracle_max_cluster_weight(g, 8)
        drv = ShardedLP(OracleBackend(g, seed=4, passes=1), g.n, 5, rank, world)
        c, moved, _ = drv.compute_clustering(mcw)
        k = 4
        part = (np.arange(g.n) % k).astype(np.uint32)
        mbw = B.oracle_max_block_weights(g, k)
        drv2 = ShardedLP(OracleBackend(g, seed=4, passes=4), g.n, 5, rank, world)
        p, bw, moved2, _ = drv2.refine(k, mbw, part)
        np.savez(out + f".{rank}.npz", c=c, p=p, bw=bw, moved=np.array(moved), moved2=np.array(moved2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_driver_matches_single_process_oracle(world, tmp_path):
    graph_seed = 5
    out = str(tmp_path / "res")
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, graph_seed, out), nprocs=world, join=True)
    g = B.oracle_rearrange(random_weights(rmat(11, 8, graph_seed), 3, max_vwgt=3))[0]
    mcw = B.oracle_max_cluster_weight(g, 8)
    expect_c, st = B.oracle_lp_cluster(g, 4, mcw, schedule=B.SYNC, return_stats=True)
    k = 4
    part = (np.arange(g.n) % k).astype(np.uint32)
    mbw = B.oracle_max_block_weights(g, k)
    rp = B.oracle_params(B.default_refine_params(), commit_passes=4)
    expect_p, expect_bw, st2 = B.oracle_lp_refine(g, 4, k, mbw, part, schedule=B.SYNC, params=rp, return_stats=True)
    for rank in range(world):
        d = np.load(out + f".{rank}.npz")
        assert np.array_equal(d["c"], expect_c), f"rank {rank}: clustering differs from the 1-process oracle"
        assert np.array_equal(d["p"], expect_p) and np.array_equal(d["bw"], expect_bw)
        assert list(d["moved"]) == list(st[0].moved[: st[0].iterations])
        assert list(d["moved2"]) == list(st2.moved[: st2.iterations])
