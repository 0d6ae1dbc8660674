"""Parity tests of the device-side cluster contraction (-m gpu): the CUDA path, called through the C ABI
(include/kaminpar_b200_contraction.h), against the CPU oracle (oracle/contraction_oracle.py) -- bit-exact
coarse CSR, weights and mapping in the canonical form -- plus the committed reference goldens and the
reference's own known-answer tests."""
import os

import numpy as np
import pytest

from kaminpar_b200 import contraction as KC
from kaminpar_b200 import lp
from kaminpar_b200.graph import CSRGraph, grid3d, random_weights, rgg2d, rmat
from oracle import bindings as B
from oracle import contraction_oracle as CO
from tests import helpers as H
from tests.test_gpu_parity import NAMES, ctx_for, get_graph

pytestmark = pytest.mark.gpu


def gpu_result(cg: KC.CoarseGraph):
    c = cg.get()
    return dict(c_n=cg.n, c_xadj=c.xadj, c_adjncy=c.adjncy, c_vwgt=c.vwgt, c_adjwgt=c.adjwgt, mapping=cg.mapping())


def oracle_of(g, cl):
    return CO.contract(g.xadj, g.adjncy, g.vwgt, g.adjwgt, cl)


def clusterings(g, seed):
    rng = np.random.default_rng(seed)
    n = g.n
    yield "identity", np.arange(n, dtype=np.uint32)
    yield "one", np.full(n, n - 1 if n else 0, np.uint32)
    yield "random", rng.integers(0, max(n, 1), n).astype(np.uint32)
    yield "coarse_random", rng.integers(0, max(n // 20, 1), n).astype(np.uint32)
    yield "runs", (np.arange(n) // 5 * 5).astype(np.uint32)
    _, mcw = ctx_for(g, 8)
    yield "lp", B.oracle_lp_cluster(g, 1, mcw, schedule=B.SYNC)


@pytest.mark.parametrize("name", NAMES)
def test_contraction_matches_oracle(name):
    g = get_graph(name)
    ctx = lp.create_default_context()
    handle = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine))
    handle.set_graph(g)
    for tag, cl in clusterings(g, 3):
        cg = KC.contract_on_handle(handle, cl)
        o = oracle_of(g, cl)
        r = gpu_result(cg)
        assert CO.equal(r, o), (name, tag)
        assert cg.stats.c_n == o["c_n"] and cg.stats.c_m == len(o["c_adjncy"])
        # projections (cluster_contraction_preprocessing.h:36-46)
        coarse = np.random.default_rng(1).integers(0, 16, o["c_n"]).astype(np.uint32)
        fine = cg.project_up(coarse)
        assert np.array_equal(fine, CO.project_up(o["mapping"], coarse))
        assert np.array_equal(cg.project_down(fine), coarse)
        cg.close()


@pytest.mark.parametrize("name", ["rgg2d_k4", "rgg16_w", "walshaw_k16", "walshaw_unsorted", "rmat13_w", "grid12",
                                  "road60", "star30000"])
def test_contraction_matches_reference_golden(name):
    g, _ = H.load_case(name)
    d = np.load(os.path.join(H.GOLDEN, f"contract_{name}.npz"))
    ref = CO.canonicalize(int(d["c_n"][0]), d["c_xadj"], d["c_adjncy"], d["c_vwgt"], d["c_adjwgt"], d["mapping"],
                          clustering=d["clustering"])
    cg = KC.contract_clustering(g, d["clustering"])
    assert CO.equal(gpu_result(cg), ref)


def test_reference_kats_on_gpu():
    from tests.test_contraction_oracle import grid2d, weighted_endpoints

    g = grid2d(2, 4)  # ContractingGridHorizontallyWorks (cluster_contraction_test.cc:107-145)
    g.vwgt = np.array([1, 2, 3, 4, 10, 20, 30, 40], np.int32)
    r = gpu_result(KC.contract_clustering(g, np.array([0, 1, 2, 3, 0, 1, 2, 3], np.uint32)))
    assert r["c_n"] == 4 and len(r["c_adjncy"]) == 6 and sorted(r["c_vwgt"]) == [11, 22, 33, 44]
    assert r["c_adjwgt"].sum() == 12 and {(11, 22), (22, 33), (33, 44)} <= weighted_endpoints(r)
    g = grid2d(2, 2)  # ContractingToSingleNodeWorks (:19-43)
    r = gpu_result(KC.contract_clustering(g, np.full(4, 2, np.uint32)))
    assert r["c_n"] == 1 and len(r["c_adjncy"]) == 0 and r["c_vwgt"][0] == 4 and list(r["c_xadj"]) == [0, 0]


def test_edge_cases():
    # empty graph
    cg = KC.contract_clustering(H.empty_graph(0), np.zeros(0, np.uint32))
    assert cg.n == 0 and cg.m == 0 and list(cg.get().xadj) == [0]
    # isolated vertices only
    g = H.empty_graph(5)
    r = gpu_result(KC.contract_clustering(g, np.array([4, 4, 1, 1, 0], np.uint32)))
    assert CO.equal(r, oracle_of(g, np.array([4, 4, 1, 1, 0], np.uint32)))
    # out-of-range cluster id -> error, no fallback
    with pytest.raises(RuntimeError, match="id >= n"):
        KC.contract_clustering(H.path_graph(10), np.full(10, 10, np.uint32))
    # a tile that spans more than kTileVerts vertices (long runs of isolated vertices between edges)
    n = 20000
    edges = [(0, n - 1), (5000, 15000), (1, 2)]
    g = H.from_edges(n, edges)
    cl = (np.arange(n) // 3).astype(np.uint32)
    assert CO.equal(gpu_result(KC.contract_clustering(g, cl)), oracle_of(g, cl))


def test_large_random_properties():
    """Size-independent properties at a size the oracle still finishes (R-MAT 18): weight conservation,
    symmetry, idempotence under the identity clustering."""
    g = random_weights(rmat(18, 16, 11), 2, max_vwgt=3, max_adjwgt=4)
    ctx, mcw = ctx_for(g, 16)
    clusterer = lp.LPClustering(ctx.coarsening, ctx.engine)
    clusterer.set_max_cluster_weight(mcw)
    cl = clusterer.compute_clustering(g)
    cg = KC.contract_on_handle(clusterer._handle, None)  # device-resident clustering
    r = gpu_result(cg)
    assert CO.equal(r, oracle_of(g, cl))
    assert r["c_vwgt"].astype(np.int64).sum() == g.total_node_weight()
    c = cg.get()
    again = gpu_result(KC.contract_clustering(c, np.arange(c.n, dtype=np.uint32)))
    assert all(np.array_equal(again[k], r[k]) for k in ("c_xadj", "c_adjncy", "c_vwgt", "c_adjwgt"))


def test_device_resident_coarsening_level():
    """LP clustering -> contraction -> LP clustering on the coarse graph without leaving the device, next to
    the oracle doing the same on the host (weighted coarse graph: vwgt + adjwgt paths of the sweeps)."""
    g = B.oracle_rearrange(rmat(15, 16, 21))[0]
    ctx, mcw = ctx_for(g, 8, seed=2)
    h0 = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine))
    h0.set_graph(g)
    _, st0 = h0.cluster(mcw, fetch=False)
    cg = KC.contract_on_handle(h0, None)
    d_xadj, d_adj, d_vw, d_ew, _ = cg.device_arrays()
    h1 = lp.LPHandle(lp._cluster_config(ctx.coarsening.clustering.lp, ctx.engine))
    h1.set_graph_device(cg.n, cg.m, d_xadj, d_adj, d_vw, d_ew)
    mcw1 = 2 * mcw
    c1, _ = h1.cluster(mcw1)
    # oracle
    cl0 = B.oracle_lp_cluster(g, 2, mcw, schedule=B.SYNC)
    o = oracle_of(g, cl0)
    assert CO.equal(gpu_result(cg), o)
    cgraph = CSRGraph(o["c_xadj"], o["c_adjncy"], o["c_vwgt"], o["c_adjwgt"])
    assert np.array_equal(c1, B.oracle_lp_cluster(cgraph, 2, mcw1, schedule=B.SYNC))
    # project the coarse clustering up: every fine vertex lands in the cluster of its coarse vertex
    assert np.array_equal(cg.project_up(c1), c1[o["mapping"]])
