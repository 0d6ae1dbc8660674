"""CPU oracle of cluster contraction -- TEST INFRASTRUCTURE ONLY.

numpy restatement of what `contract_clustering(graph, clustering, con_ctx)` computes
(kaminpar-shm/coarsening/contraction/cluster_contraction.cc:22-50):

  * fine -> coarse mapping: coarse ids are the ranks of the used cluster (leader) ids
    (cluster_contraction_preprocessing.cc:17-51: flag `leader_mapping[clustering[u]]`, prefix sum,
    `mapping[u] = leader_mapping[clustering[u]] - 1`);
  * coarse node weight = sum of the member weights (unbuffered_cluster_contraction.cc:271-289,
    `c_u_weight += graph.node_weight(u)`);
  * coarse edge (c_u, c_v), c_u != c_v, weight = sum over fine edges (u, v) with mapping[u] = c_u,
    mapping[v] = c_v (`local_edge_collector[c_v] += w`, :280-285); edges inside a cluster vanish.

The reference then RENUMBERS the coarse vertices in the order its threads finish them
(`remapping`, unbuffered_cluster_contraction.cc:172-179, 558-572) and emits each adjacency list in
hash-map insertion order; both are scheduling artefacts (they differ from run to run with more than
one thread, and the reference's own tests only use order-free matchers,
tests/shm/coarsening/cluster_contraction_test.cc). The canonical form used for parity is therefore:
coarse ids in leader order, adjacency lists sorted by target. `canonicalize()` brings any valid
output (the reference's, in particular) into that form.

Pinned against the unmodified reference (oracle/_ref, all three contraction algorithms) by
tests/test_contraction_oracle.py and the fixtures in tests/golden/contract_*.npz.
"""
from __future__ import annotations

import numpy as np


def contract(xadj, adjncy, vwgt, adjwgt, clustering):
    """Returns dict(c_n, c_xadj, c_adjncy, c_vwgt, c_adjwgt, mapping) in canonical form."""
    xadj = np.asarray(xadj, np.int64)
    adjncy = np.asarray(adjncy, np.int64)
    n = len(xadj) - 1
    m = len(adjncy)
    clustering = np.asarray(clustering, np.int64)
    if n == 0:
        z32 = np.zeros(0, np.uint32)
        return dict(c_n=0, c_xadj=np.zeros(1, np.uint32), c_adjncy=z32, c_vwgt=np.zeros(0, np.int32),
                    c_adjwgt=np.zeros(0, np.int32), mapping=z32)
    # cluster_contraction_preprocessing.cc:27-33
    leader = np.zeros(n, np.int64)
    leader[clustering] = 1
    leader = np.cumsum(leader)
    c_n = int(leader[n - 1])
    mapping = leader[clustering] - 1  # :44-46
    vw = np.ones(n, np.int64) if vwgt is None else np.asarray(vwgt, np.int64)
    c_vwgt = np.bincount(mapping, weights=vw, minlength=c_n).astype(np.int64)
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(xadj))
    cu = mapping[src]
    cv = mapping[adjncy]
    ew = np.ones(m, np.int64) if adjwgt is None else np.asarray(adjwgt, np.int64)
    keep = cu != cv  # unbuffered_cluster_contraction.cc:282
    key = cu[keep] * c_n + cv[keep]
    uniq, inv = np.unique(key, return_inverse=True)
    c_adjwgt = np.bincount(inv, weights=ew[keep], minlength=len(uniq)).astype(np.int64)
    c_src = uniq // c_n
    c_adjncy = uniq % c_n
    c_xadj = np.zeros(c_n + 1, np.int64)
    np.add.at(c_xadj, c_src + 1, 1)
    c_xadj = np.cumsum(c_xadj)
    return dict(c_n=c_n, c_xadj=c_xadj.astype(np.uint32), c_adjncy=c_adjncy.astype(np.uint32),
                c_vwgt=c_vwgt.astype(np.int32), c_adjwgt=c_adjwgt.astype(np.int32), mapping=mapping.astype(np.uint32))


def canonicalize(c_n, c_xadj, c_adjncy, c_vwgt, c_adjwgt, mapping, clustering):
    """Relabel a contraction result so that coarse ids follow leader order and every adjacency list is
    sorted by target. `clustering` is the fine clustering the result was computed from: the coarse
    vertex of leader id L gets the rank of L among the used leader ids."""
    c_xadj = np.asarray(c_xadj, np.int64)
    c_adjncy = np.asarray(c_adjncy, np.int64)
    mapping = np.asarray(mapping, np.int64)
    clustering = np.asarray(clustering, np.int64)
    n = len(mapping)
    if n == 0:
        return contract(np.zeros(1), np.zeros(0), None, None, np.zeros(0))
    leader = np.zeros(n, np.int64)
    leader[clustering] = 1
    canon_of_fine = np.cumsum(leader)[clustering] - 1
    # the old coarse id of every fine vertex is mapping[u]; all members of a cluster must agree
    perm = np.full(c_n, -1, np.int64)  # old coarse id -> canonical id
    perm[mapping] = canon_of_fine
    assert (perm >= 0).all() and len(np.unique(perm)) == c_n, "mapping is not a bijection onto the clusters"
    assert (perm[mapping] == canon_of_fine).all(), "vertices of one cluster map to different coarse vertices"
    deg = np.diff(c_xadj)
    src = np.repeat(np.arange(c_n, dtype=np.int64), deg)
    key = perm[src] * c_n + perm[c_adjncy[: c_xadj[c_n]]]
    order = np.argsort(key, kind="stable")
    key = key[order]
    assert len(np.unique(key)) == len(key), "duplicate coarse edge"
    new_xadj = np.zeros(c_n + 1, np.int64)
    np.add.at(new_xadj, key // c_n + 1, 1)
    inv = np.empty(c_n, np.int64)
    inv[perm] = np.arange(c_n)
    return dict(c_n=c_n, c_xadj=np.cumsum(new_xadj).astype(np.uint32), c_adjncy=(key % c_n).astype(np.uint32),
                c_vwgt=np.asarray(c_vwgt, np.int32)[inv].copy(),
                c_adjwgt=np.asarray(c_adjwgt, np.int32)[: c_xadj[c_n]][order].copy(),
                mapping=perm[mapping].astype(np.uint32))


def equal(a, b) -> bool:
    return (a["c_n"] == b["c_n"] and all(np.array_equal(a[k], b[k])
                                         for k in ("c_xadj", "c_adjncy", "c_vwgt", "c_adjwgt", "mapping")))


def project_up(mapping, coarse):  # CoarseGraph::project_up (cluster_contraction.h:28): fine[u] = coarse[mapping[u]]
    return np.asarray(coarse)[np.asarray(mapping, np.int64)]


def project_down(mapping, fine, c_n):  # CoarseGraph::project_down (:30): coarse[mapping[u]] = fine[u]
    out = np.zeros(c_n, np.asarray(fine).dtype)
    out[np.asarray(mapping, np.int64)] = fine
    return out
