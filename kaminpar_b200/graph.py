"""Host-side CSR graph container, METIS reader and synthetic generators.

Layout mirrors the reference's ``CSRGraph`` (kaminpar-shm/datastructures/csr_graph.h:35-482):
``xadj[n+1]`` (EdgeID=uint32), ``adjncy[m]`` (NodeID=uint32), optional ``vwgt[n]`` / ``adjwgt[m]``
(int32); *absent weight arrays mean unit weights* (csr_graph.cc:82-97).

The generators implement the synthetic inputs of SURVEY.md §8(d) / BASELINE.md §3 (the reference
ships no generator; KaGen is an un-vendored dependency). They are written with torch ops so they
run on the CPU for tests and on the GPU for benchmark-sized inputs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class CSRGraph:
    xadj: np.ndarray            # uint32 [n+1]
    adjncy: np.ndarray          # uint32 [m]
    vwgt: Optional[np.ndarray] = None    # int32 [n] or None (unit)
    adjwgt: Optional[np.ndarray] = None  # int32 [m] or None (unit)
    sorted: bool = False        # degree-bucket sorted (reference: CSRGraph::sorted())
    buckets: Optional[np.ndarray] = None  # uint32 [34] bucket prefix array when sorted

    def __post_init__(self):
        self.xadj = np.ascontiguousarray(self.xadj, dtype=np.uint32)
        self.adjncy = np.ascontiguousarray(self.adjncy, dtype=np.uint32)
        if self.vwgt is not None:
            self.vwgt = np.ascontiguousarray(self.vwgt, dtype=np.int32)
        if self.adjwgt is not None:
            self.adjwgt = np.ascontiguousarray(self.adjwgt, dtype=np.int32)

    @property
    def n(self) -> int:
        return int(self.xadj.shape[0] - 1)

    @property
    def m(self) -> int:
        return int(self.adjncy.shape[0])

    def degrees(self) -> np.ndarray:
        return np.diff(self.xadj.astype(np.int64)).astype(np.uint32)

    def total_node_weight(self) -> int:
        return int(self.n if self.vwgt is None else self.vwgt.astype(np.int64).sum())

    def max_node_weight(self) -> int:
        return 1 if self.vwgt is None or self.n == 0 else int(self.vwgt.max())


# --------------------------------------------------------------------------------------------
# METIS text format (reference: docs/graph_file_format.md, kaminpar-io/metis_parser.cc:158-298)
# --------------------------------------------------------------------------------------------
def read_metis(path: str) -> CSRGraph:
    with open(path, "r") as f:
        lines = [ln for ln in f.read().split("\n")]
    it = iter(lines)
    header = None
    for ln in it:
        s = ln.strip()
        if not s or s.startswith("%"):
            continue
        header = s.split()
        break
    if header is None:
        raise ValueError("empty METIS file")
    n, m_undirected = int(header[0]), int(header[1])
    fmt = int(header[2]) if len(header) > 2 else 0
    has_ew = (fmt % 10) == 1
    has_vw = ((fmt // 10) % 10) == 1
    xadj = np.zeros(n + 1, dtype=np.int64)
    adj, ew, vw = [], [], []
    u = 0
    for ln in it:
        if u >= n:
            break
        s = ln.strip()
        if s.startswith("%"):
            continue
        toks = [int(t) for t in s.split()]
        if has_vw:
            vw.append(toks[0])
            toks = toks[1:]
        if has_ew:
            adj.extend(t - 1 for t in toks[0::2])
            ew.extend(toks[1::2])
            xadj[u + 1] = xadj[u] + len(toks) // 2
        else:
            adj.extend(t - 1 for t in toks)
            xadj[u + 1] = xadj[u] + len(toks)
        u += 1
    while u < n:  # trailing isolated nodes with missing lines
        xadj[u + 1] = xadj[u]
        if has_vw:
            vw.append(1)
        u += 1
    g = CSRGraph(
        xadj=xadj.astype(np.uint32),
        adjncy=np.asarray(adj, dtype=np.uint32),
        vwgt=np.asarray(vw, dtype=np.int32) if has_vw else None,
        adjwgt=np.asarray(ew, dtype=np.int32) if has_ew else None,
    )
    if g.m != 2 * m_undirected:
        raise ValueError(f"METIS header says {m_undirected} edges, file holds {g.m} directed entries")
    return g


def write_metis(g: CSRGraph, path: str) -> None:
    fmt = (10 if g.vwgt is not None else 0) + (1 if g.adjwgt is not None else 0)
    with open(path, "w") as f:
        f.write(f"{g.n} {g.m // 2}" + (f" {fmt:02d}" if fmt else "") + "\n")
        for u in range(g.n):
            parts = []
            if g.vwgt is not None:
                parts.append(str(int(g.vwgt[u])))
            for e in range(int(g.xadj[u]), int(g.xadj[u + 1])):
                parts.append(str(int(g.adjncy[e]) + 1))
                if g.adjwgt is not None:
                    parts.append(str(int(g.adjwgt[e])))
            f.write(" ".join(parts) + "\n")


# --------------------------------------------------------------------------------------------
# Edge list -> symmetric CSR (torch: runs on CPU or CUDA)
# --------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------
# ParHIP binary format + partition files (reference: docs/graph_file_format.md,
# kaminpar-io/parhip_parser.cc:42-93, kaminpar-io/kaminpar_io.cc:58-75)
# --------------------------------------------------------------------------------------------
def read_parhip(path: str) -> CSRGraph:
    """24-byte header (version bit-field, n, m), offsets (file addresses of each adjacency list), adjacency lists,
    optional node / edge weights. Version bits (0 = present / 64-bit): 1 edge weights, 2 node weights, 4 edge-id
    width, 8 node-id width, 16 node-weight width, 32 edge-weight width (parhip_parser.cc:80-93)."""
    raw = np.fromfile(path, dtype=np.uint8)
    version, n, m = (int(x) for x in raw[:24].view("<u8"))
    has_ew, has_nw = (version & 1) == 0, (version & 2) == 0
    eid = np.dtype("<u8") if (version & 4) == 0 else np.dtype("<u4")
    nid = np.dtype("<u8") if (version & 8) == 0 else np.dtype("<u4")
    nwt = np.dtype("<i8") if (version & 16) == 0 else np.dtype("<i4")
    ewt = np.dtype("<i8") if (version & 32) == 0 else np.dtype("<i4")
    pos = 24
    offsets = raw[pos:pos + (n + 1) * eid.itemsize].view(eid).astype(np.int64)
    pos += (n + 1) * eid.itemsize
    base = 24 + (n + 1) * eid.itemsize  # _nodes_offset_base
    xadj = (offsets - base) // nid.itemsize  # map_edge_offset (parhip_parser.cc:112-114)
    adjncy = raw[pos:pos + m * nid.itemsize].view(nid)
    pos += m * nid.itemsize
    vwgt = adjwgt = None
    if has_nw:
        vwgt = raw[pos:pos + n * nwt.itemsize].view(nwt).astype(np.int32)
        pos += n * nwt.itemsize
    if has_ew:
        adjwgt = raw[pos:pos + m * ewt.itemsize].view(ewt).astype(np.int32)
    if int(xadj[-1]) != m or (m and int(adjncy.max()) >= n):
        raise ValueError(f"{path}: inconsistent ParHIP file")
    return CSRGraph(xadj.astype(np.uint32), adjncy.astype(np.uint32), vwgt, adjwgt)


def write_parhip(g: CSRGraph, path: str) -> None:
    """32-bit ids and weights (the default build's widths, kaminpar.h:32-57)."""
    has_nw, has_ew = g.vwgt is not None, g.adjwgt is not None
    version = (0 if has_ew else 1) | (0 if has_nw else 2) | 4 | 8 | 16 | 32
    base = 24 + (g.n + 1) * 4
    with open(path, "wb") as f:
        np.array([version, g.n, g.m], "<u8").tofile(f)
        (g.xadj.astype(np.int64) * 4 + base).astype("<u4").tofile(f)
        g.adjncy.astype("<u4").tofile(f)
        if has_nw:
            g.vwgt.astype("<i4").tofile(f)
        if has_ew:
            g.adjwgt.astype("<i4").tofile(f)


def write_partition(path: str, partition) -> None:
    """One block id per line (kaminpar_io.cc:58-63)."""
    np.savetxt(path, np.asarray(partition, dtype=np.int64), fmt="%d")


def read_partition(path: str) -> np.ndarray:
    """kaminpar_io.cc:65-75"""
    return np.loadtxt(path, dtype=np.int64, ndmin=1).astype(np.uint32)


def _csr_from_pairs_torch(n: int, src, dst, device):
    """Symmetrise, drop self-loops and duplicates, return (xadj:int64[n+1], adjncy:int64[m])."""
    import torch

    keep = src != dst
    src, dst = src[keep], dst[keep]
    a = torch.cat([src, dst])
    b = torch.cat([dst, src])
    del src, dst
    key = a * n + b
    del a, b
    key = torch.unique(key)  # sorted, deduplicated
    u = torch.div(key, n, rounding_mode="floor")
    v = key - u * n
    del key
    deg = torch.bincount(u, minlength=n)
    xadj = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=xadj[1:])
    return xadj, v


def _to_graph(xadj, adjncy) -> CSRGraph:
    return CSRGraph(
        xadj=xadj.to("cpu").numpy().astype(np.uint32), adjncy=adjncy.to("cpu").numpy().astype(np.uint32)
    )


def rmat_edges_torch(scale: int, edge_factor: int, seed: int, device="cpu", a=0.57, b=0.19, c=0.19):
    """Graph500-style R-MAT sampler (SURVEY §8d input 2/4): a=.57 b=.19 c=.19 d=.05."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    num = edge_factor << scale
    src = torch.zeros(num, dtype=torch.int64, device=device)
    dst = torch.zeros(num, dtype=torch.int64, device=device)
    ab, abc = a + b, a + b + c
    for _ in range(scale):
        r = torch.rand(num, generator=gen, device=device)
        src_bit = (r >= ab).to(torch.int64)                      # quadrants c, d -> lower half rows
        dst_bit = (((r >= a) & (r < ab)) | (r >= abc)).to(torch.int64)  # quadrants b, d
        src = (src << 1) | src_bit
        dst = (dst << 1) | dst_bit
    return src, dst


def rmat(scale: int, edge_factor: int = 16, seed: int = 1, device="cpu") -> CSRGraph:
    n = 1 << scale
    src, dst = rmat_edges_torch(scale, edge_factor, seed, device)
    xadj, adj = _csr_from_pairs_torch(n, src, dst, device)
    return _to_graph(xadj, adj)


def grid3d(nx: int, ny: Optional[int] = None, nz: Optional[int] = None) -> CSRGraph:
    """6-neighbour 3-D grid, lexicographic ids (SURVEY §8d input 3)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = nx * ny * nz
    ids = np.arange(n, dtype=np.int64)
    x = ids % nx
    y = (ids // nx) % ny
    z = ids // (nx * ny)
    # neighbours in ascending id order: -z, -y, -x, +x, +y, +z
    cand = [
        (z > 0, -nx * ny), (y > 0, -nx), (x > 0, -1), (x < nx - 1, 1), (y < ny - 1, nx), (z < nz - 1, nx * ny),
    ]
    deg = np.zeros(n, dtype=np.int64)
    for mask, _ in cand:
        deg += mask
    xadj = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=xadj[1:])
    adjncy = np.empty(int(xadj[-1]), dtype=np.uint32)
    pos = xadj[:-1].copy()
    for mask, off in cand:
        idx = ids[mask]
        adjncy[pos[idx]] = (idx + off).astype(np.uint32)
        pos[idx] += 1
    return CSRGraph(xadj=xadj.astype(np.uint32), adjncy=adjncy)


def grid3d_torch(nx: int, device="cpu"):
    """grid3d on a torch device (benchmark sizes: 512^3 = 134 M vertices): returns int64 (xadj, adjncy) tensors with
    the same neighbour order as grid3d (-z, -y, -x, +x, +y, +z)."""
    import torch

    n = nx * nx * nx
    ids = torch.arange(n, dtype=torch.int64, device=device)
    x = ids % nx
    y = (ids // nx) % nx
    z = ids // (nx * nx)
    cand = [(z > 0, -nx * nx), (y > 0, -nx), (x > 0, -1), (x < nx - 1, 1), (y < nx - 1, nx), (z < nx - 1, nx * nx)]
    del x, y, z
    deg = torch.zeros(n, dtype=torch.int64, device=device)
    for mask, _ in cand:
        deg += mask
    xadj = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=xadj[1:])
    del deg
    adjncy = torch.empty(int(xadj[-1].item()), dtype=torch.int64, device=device)
    pos = xadj[:-1].clone()
    for mask, off in cand:
        adjncy[pos[mask]] = ids[mask] + off
        pos += mask
    return xadj, adjncy


def rgg2d(n: int, seed: int = 1, radius_factor: float = 0.55, device="cpu") -> CSRGraph:
    """Random geometric graph in the unit square, r = radius_factor*sqrt(ln n / n) (input 6)."""
    import math

    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    r = radius_factor * math.sqrt(math.log(n) / n)
    cells = max(1, int(1.0 / r))
    pts = torch.rand(n, 2, generator=gen, device=device, dtype=torch.float64)
    cx = torch.clamp((pts[:, 0] * cells).to(torch.int64), max=cells - 1)
    cy = torch.clamp((pts[:, 1] * cells).to(torch.int64), max=cells - 1)
    cell = cy * cells + cx
    order = torch.argsort(cell, stable=True)
    pts, cell, cx, cy = pts[order], cell[order], cx[order], cy[order]
    counts = torch.bincount(cell, minlength=cells * cells)
    start = torch.zeros(cells * cells + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=start[1:])
    max_cnt = int(counts.max().item())
    srcs, dsts = [], []
    ids = torch.arange(n, device=device)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            nx_, ny_ = cx + dx, cy + dy
            ok = (nx_ >= 0) & (nx_ < cells) & (ny_ >= 0) & (ny_ < cells)
            ncell = torch.where(ok, ny_ * cells + nx_, torch.zeros_like(cx))
            nstart = start[ncell]
            ncount = torch.where(ok, counts[ncell], torch.zeros_like(cx))
            for j in range(max_cnt):
                valid = j < ncount
                if not bool(valid.any()):
                    break
                v = torch.clamp(nstart + j, max=n - 1)
                d = pts - pts[v]
                close = valid & ((d * d).sum(1) <= r * r) & (v > ids)
                srcs.append(ids[close])
                dsts.append(v[close])
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.int64, device=device)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.int64, device=device)
    xadj, adj = _csr_from_pairs_torch(n, src, dst, device)
    return _to_graph(xadj, adj)


def road_like(side: int, seed: int = 1, delete_frac: float = 0.2, subdivide_frac: float = 0.0,
              device="cpu") -> CSRGraph:
    """Planar low-degree stand-in for a road network (input 5): side x side 4-neighbour grid with
    ``delete_frac`` of the edges removed and ``subdivide_frac`` of the remaining edges subdivided
    by one degree-2 vertex."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n0 = side * side
    ids = torch.arange(n0, device=device)
    x, y = ids % side, ids // side
    h_src = ids[x < side - 1]
    v_src = ids[y < side - 1]
    src = torch.cat([h_src, v_src])
    dst = torch.cat([h_src + 1, v_src + side])
    keep = torch.rand(src.numel(), generator=gen, device=device) >= delete_frac
    src, dst = src[keep], dst[keep]
    n = n0
    if subdivide_frac > 0:
        sub = torch.rand(src.numel(), generator=gen, device=device) < subdivide_frac
        k = int(sub.sum().item())
        mid = torch.arange(n0, n0 + k, device=device)
        s2 = torch.cat([src[~sub], src[sub], mid])
        d2 = torch.cat([dst[~sub], mid, dst[sub]])
        src, dst, n = s2, d2, n0 + k
    xadj, adj = _csr_from_pairs_torch(n, src, dst, device)
    return _to_graph(xadj, adj)


def random_weights(g: CSRGraph, seed: int, max_vwgt: int = 0, max_adjwgt: int = 0) -> CSRGraph:
    """Attach symmetric random edge weights / random node weights (test helper)."""
    rng = np.random.default_rng(seed)
    vwgt = rng.integers(1, max_vwgt + 1, size=g.n).astype(np.int32) if max_vwgt > 0 else None
    adjwgt = None
    if max_adjwgt > 0:
        src = np.repeat(np.arange(g.n, dtype=np.int64), np.diff(g.xadj.astype(np.int64)))
        dst = g.adjncy.astype(np.int64)
        lo, hi = np.minimum(src, dst), np.maximum(src, dst)
        h = (lo * 0x9E3779B97F4A7C15 + hi * 0xC2B2AE3D27D4EB4F + seed) & 0xFFFFFFFFFFFFFFFF if False else None
        # symmetric hash of the undirected pair
        key = (lo * np.int64(1000003) + hi * np.int64(998244353) + np.int64(seed)) & np.int64(0x7FFFFFFF)
        adjwgt = (key % max_adjwgt + 1).astype(np.int32)
    return CSRGraph(xadj=g.xadj, adjncy=g.adjncy, vwgt=vwgt, adjwgt=adjwgt, sorted=g.sorted, buckets=g.buckets)


# --------------------------------------------------------------------------------------------
# Degree-bucket rearrangement (what KaMinPar::compute_partition applies before any LP runs:
# kaminpar-shm/kaminpar.cc:369-396, graphutils/permutator.h:28-208, csr_graph.cc:150-174)
# --------------------------------------------------------------------------------------------
def rearrange_by_degree_buckets_torch(xadj, adjncy, remove_isolated=True):
    """torch (CPU or CUDA) version for benchmark-sized inputs: stable sort of the vertices by degree
    bucket (0 if d == 0 else floor(log2 d) + 1, isolated vertices last), neighbours relabelled and
    every adjacency list reversed, isolated vertices cut off the end. Unit weights only.
    Returns (xadj:int64[n'+1], adjncy:int64[m], old_to_new:int64[n])."""
    import torch

    n = xadj.numel() - 1
    deg = xadj[1:] - xadj[:-1]
    bucket = torch.where(deg > 0, torch.floor(torch.log2(deg.clamp(min=1).double())).long() + 1,
                         torch.full_like(deg, 40))
    # exact integer floor(log2): fix potential fp rounding at powers of two
    bucket = torch.where((deg > 0) & ((1 << (bucket - 1).clamp(min=0, max=62)) > deg), bucket - 1, bucket)
    new_to_old = torch.argsort(bucket, stable=True)
    old_to_new = torch.empty_like(new_to_old)
    old_to_new[new_to_old] = torch.arange(n, device=xadj.device)
    new_deg = deg[new_to_old]
    new_xadj = torch.zeros(n + 1, dtype=torch.int64, device=xadj.device)
    torch.cumsum(new_deg, 0, out=new_xadj[1:])
    m = adjncy.numel()
    src_old = torch.repeat_interleave(torch.arange(n, device=xadj.device), deg)
    e = torch.arange(m, device=xadj.device)
    u_new = old_to_new[src_old]
    pos = new_xadj[u_new + 1] - 1 - (e - xadj[src_old])
    del src_old, e, u_new
    new_adj = torch.empty_like(adjncy)
    new_adj[pos] = old_to_new[adjncy]
    del pos
    n_lp = n
    if remove_isolated:
        n_lp = int((deg > 0).sum().item())
    return new_xadj[: n_lp + 1].contiguous(), new_adj, old_to_new
