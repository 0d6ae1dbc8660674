"""Sharded multi-GPU driver of the LP engine: one process per GPU, `torch.distributed` (NCCL over
NVLink / NVSwitch) for the exchange.

Design (DESIGN.md "multi-GPU"): every rank holds a replica of the CSR graph and of the label /
weight / active state; the *vertex frontier* is sharded -- each rank sweeps its share of every
(degree group, sub-round) work list. Per sub-round:

    sweep (rank's share)  ->  all_gather of the packed proposals  ->  identical commit on every rank

The commit rule is order-independent, so the replicas stay bit-identical and the result equals the
single-GPU (and the CPU oracle's `sync`) result for any world size. This is the role of the
reference's distributed twin (kaminpar-dist/refinement/lp/lp_refiner.cc:119-222: local
`perform_iteration(from, to)` per chunk, then label exchange; MPI call sites in SURVEY.md §2.2),
with NCCL all-gather instead of MPI sparse all-to-all and no ghost-vertex bookkeeping.

The driver is backend-agnostic: `CudaBackend` drives the stepping C ABI
(kmp_lp_step_* in include/kaminpar_b200_lp.h); tests plug in a CPU backend to exercise the exchange
logic under the `gloo` backend.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import lp


class CudaBackend:
    """Stepping C ABI on this rank's GPU. Buffers are torch CUDA tensors (int32 words)."""

    def __init__(self, handle: lp.LPHandle, device):
        import torch

        self.h = handle
        self.lib = lp.load_library()
        self.device = device
        self.torch = torch

    def set_shard(self, rank, world):
        lp._check(self.lib.kmp_lp_set_shard(self.h._h, C.c_uint32(rank), C.c_uint32(world)))
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        lp._check(self.lib.kmp_lp_set_stream(self.h._h, C.c_void_p(stream)))

    def num_subrounds(self):
        return int(self.lib.kmp_lp_num_subrounds(self.h._h))

    def subround_cap(self, sg):
        cap, size = C.c_uint32(0), C.c_uint32(0)
        lp._check(self.lib.kmp_lp_subround_cap(self.h._h, C.c_uint32(sg), C.byref(cap), C.byref(size)))
        return cap.value, size.value

    def alloc(self, words):
        return self.torch.empty(words, dtype=self.torch.int32, device=self.device)

    def begin_cluster(self, max_cluster_weight, communities):
        comm = None if communities is None else np.ascontiguousarray(communities, np.uint32)
        lp._check(self.lib.kmp_lp_step_begin_cluster(self.h._h, C.c_int32(int(max_cluster_weight)), lp._ptr(comm)))

    def begin_refine(self, k, max_bw, min_bw, communities, partition):
        mbw = np.ascontiguousarray(max_bw, np.int32)
        mnw = None if min_bw is None else np.ascontiguousarray(min_bw, np.int32)
        comm = None if communities is None else np.ascontiguousarray(communities, np.uint32)
        part = np.ascontiguousarray(partition, np.uint32)
        lp._check(self.lib.kmp_lp_step_begin_refine(self.h._h, C.c_uint32(int(k)), lp._ptr(mbw), lp._ptr(mnw),
                                                    lp._ptr(comm), lp._ptr(part)))

    def begin_iteration(self):
        lp._check(self.lib.kmp_lp_step_begin_iteration(self.h._h))

    def sweep(self, it, sg, send):
        lp._check(self.lib.kmp_lp_step_sweep(self.h._h, C.c_uint32(it), C.c_uint32(sg), C.c_void_p(send.data_ptr())))

    def commit(self, it, sg, gathered):
        lp._check(self.lib.kmp_lp_step_commit(self.h._h, C.c_uint32(it), C.c_uint32(sg),
                                              C.c_void_p(gathered.data_ptr())))

    def end_iteration(self):
        moved = C.c_uint32(0)
        lp._check(self.lib.kmp_lp_step_end_iteration(self.h._h, C.byref(moved)))
        return moved.value

    def favored_export(self, buf):
        lp._check(self.lib.kmp_lp_step_favored_export(self.h._h, C.c_void_p(buf.data_ptr())))

    def favored_import(self, buf):
        lp._check(self.lib.kmp_lp_step_favored_import(self.h._h, C.c_void_p(buf.data_ptr())))

    def finish(self, n, k=None, fetch=True):
        stats = lp.KmpStats()
        out = np.empty(n, np.uint32) if fetch else None
        bw = np.zeros(k, np.int32) if k else None
        lp._check(self.lib.kmp_lp_step_finish(self.h._h, lp._ptr(out), lp._ptr(bw), C.byref(stats)))
        return out, bw, stats


class ShardedLP:
    """Runs compute_clustering / refine with the vertex frontier sharded over the ranks of a
    process group. All ranks must call the same methods with the same arguments."""

    def __init__(self, backend, n, num_iterations, rank=0, world=1, group=None, two_hop=True):
        self.b = backend
        self.n = n
        self.num_iterations = num_iterations
        self.rank, self.world, self.group = rank, world, group
        self.two_hop = two_hop
        self.b.set_shard(rank, world)
        self._bufs = {}
        self.exchanged_words = 0

    def _exchange(self, sg, it):
        import torch.distributed as dist

        cap, size = self.b.subround_cap(sg)
        if size == 0:
            return
        words = 4 + 2 * cap
        if words not in self._bufs:
            self._bufs[words] = (self.b.alloc(words), self.b.alloc(words * self.world))
        send, recv = self._bufs[words]
        self.b.sweep(it, sg, send)
        if self.world > 1:
            dist.all_gather_into_tensor(recv, send, group=self.group)
            self.exchanged_words += words * self.world
        else:
            recv = send
        self.b.commit(it, sg, recv)

    def _iterations(self):
        moved_per_round = []
        max_it = self.num_iterations if self.num_iterations > 0 else (1 << 62)
        it = 0
        while it < max_it:
            self.b.begin_iteration()
            for sg in range(self.b.num_subrounds()):
                self._exchange(sg, it)
            moved = self.b.end_iteration()  # identical on every rank (replicated commit)
            moved_per_round.append(moved)
            it += 1
            if moved == 0:
                break
        return moved_per_round

    def compute_clustering(self, max_cluster_weight, communities=None, fetch=True):
        import torch.distributed as dist

        self.b.begin_cluster(max_cluster_weight, communities)
        moved = self._iterations()
        if self.world > 1 and self.two_hop:
            buf = self.b.alloc(self.n)
            self.b.favored_export(buf)
            dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=self.group)
            self.b.favored_import(buf)
        out, _, stats = self.b.finish(self.n, fetch=fetch)
        return out, moved, stats

    def refine(self, k, max_block_weights, partition, min_block_weights=None, communities=None):
        self.b.begin_refine(k, max_block_weights, min_block_weights, communities, partition)
        moved = self._iterations()
        out, bw, stats = self.b.finish(self.n, k=k)
        return out, bw, moved, stats


def make_cuda_sharded_clusterer(ctx: lp.Context, n, m, d_xadj_ptr, d_adjncy_ptr, rank, world, device, group=None):
    """Convenience: clusterer over device-resident CSR arrays (pointers), sharded over `world` ranks."""
    handle = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine))
    handle.set_graph_device(n, m, d_xadj_ptr, d_adjncy_ptr)
    backend = CudaBackend(handle, device)
    return ShardedLP(backend, n, ctx.coarsening.clustering.lp.num_iterations, rank, world, group), handle
