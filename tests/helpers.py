"""Shared test helpers (graph families mirror tests/shm/graph_factories.h:16-183 of the reference)."""
import glob
import os

import numpy as np

from kaminpar_b200.graph import CSRGraph

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_graph(name: str) -> CSRGraph:
    d = np.load(os.path.join(GOLDEN, f"graph_{name}.npz"))
    return CSRGraph(d["xadj"], d["adjncy"], d["vwgt"] if "vwgt" in d else None, d["adjwgt"] if "adjwgt" in d else None)


def golden_cases():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


def load_case(name: str):
    d = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    g = CSRGraph(d["xadj"], d["adjncy"], d["vwgt"] if "vwgt" in d else None, d["adjwgt"] if "adjwgt" in d else None,
                 sorted=bool(d["sorted"][0]), buckets=d["buckets"] if "buckets" in d else None)
    return g, d


def from_edges(n, edges, vwgt=None, ew=None) -> CSRGraph:
    """Undirected edge list -> symmetric CSR (tests/shm/graph_builder.h)."""
    adj = [[] for _ in range(n)]
    wts = [[] for _ in range(n)]
    for i, (u, v) in enumerate(edges):
        w = 1 if ew is None else ew[i]
        adj[u].append(v)
        wts[u].append(w)
        adj[v].append(u)
        wts[v].append(w)
    xadj = np.zeros(n + 1, np.uint32)
    for u in range(n):
        xadj[u + 1] = xadj[u] + len(adj[u])
    adjncy = np.array([v for a in adj for v in a], np.uint32)
    adjwgt = None if ew is None else np.array([w for a in wts for w in a], np.int32)
    return CSRGraph(xadj, adjncy, None if vwgt is None else np.asarray(vwgt, np.int32), adjwgt)


def empty_graph(n=0):
    return CSRGraph(np.zeros(n + 1, np.uint32), np.zeros(0, np.uint32))


def path_graph(n):
    return from_edges(n, [(i, i + 1) for i in range(n - 1)])


def star_graph(leaves):
    return from_edges(leaves + 1, [(0, i + 1) for i in range(leaves)])


def complete_graph(n):
    return from_edges(n, [(i, j) for i in range(n) for j in range(i + 1, n)])


def complete_bipartite(a, b):
    return from_edges(a + b, [(i, a + j) for i in range(a) for j in range(b)])


def grid2d(rows, cols):
    e = []
    for r in range(rows):
        for c in range(cols):
            u = r * cols + c
            if c + 1 < cols:
                e.append((u, u + 1))
            if r + 1 < rows:
                e.append((u, u + cols))
    return from_edges(rows * cols, e)


def matching_graph(pairs):
    return from_edges(2 * pairs, [(2 * i, 2 * i + 1) for i in range(pairs)])


def big_star(n: int) -> CSRGraph:
    xadj = np.zeros(n + 1, np.int64)
    xadj[1] = n - 1
    xadj[2:] = n - 1 + np.arange(1, n)
    adj = np.concatenate([np.arange(1, n), np.zeros(n - 1, np.int64)])
    return CSRGraph(xadj.astype(np.uint32), adj.astype(np.uint32))


def block_weights(g: CSRGraph, part, k):
    w = np.ones(g.n, np.int64) if g.vwgt is None else g.vwgt.astype(np.int64)
    return np.bincount(part, weights=w, minlength=k).astype(np.int64)


def cluster_weights_ok(g: CSRGraph, clustering, max_w):
    w = np.ones(g.n, np.int64) if g.vwgt is None else g.vwgt.astype(np.int64)
    cw = np.bincount(clustering, weights=w, minlength=g.n)
    mx = max(int(max_w), int(w.max()) if g.n else 0)
    return bool((cw <= mx).all())
