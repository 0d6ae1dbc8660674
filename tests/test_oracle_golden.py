"""Pins the oracle: its `seq` schedule must reproduce, bit for bit, the outputs of the UNMODIFIED
reference (LPClustering / LabelPropagationRefiner at one thread) stored under tests/golden/ref_*.npz
by tests/golden/make_golden.py. Also checks the graph utilities and scalar helpers against values the
reference produced, and the properties the reference's own tests assert
(tests/endtoend/shm_endtoend_test.cc:142-247, bindings/python/tests/test_pykaminpar.py:78-104)."""
import numpy as np
import pytest

from oracle import bindings as B
from tests import helpers as H


def _params(d, tag, default):
    p = default
    if tag in d:
        it, impl, tie, ths, iso = [int(x) for x in d[tag]]
        p.num_iterations, p.impl, p.tie_breaking = it, impl, tie
        p.two_hop_strategy, p.isolated_nodes_strategy = ths, iso
    return p


@pytest.mark.parametrize("name", H.golden_cases())
def test_seq_oracle_reproduces_reference(name):
    g, d = H.load_case(name)
    k = int(d["k"][0])
    mcw = int(d["max_cluster_weight"][0])
    mbw = d["max_block_weights"]
    num_calls = int(d["num_calls"][0])
    cp = _params(d, "cparams", B.default_cluster_params())
    rp = _params(d, "rparams", B.default_refine_params())
    assert B.oracle_max_cluster_weight(g, k) == mcw
    assert np.array_equal(B.oracle_max_block_weights(g, k), mbw)
    for seed in d["seeds"]:
        seed = int(seed)
        c = B.oracle_lp_cluster(g, seed, mcw, schedule=B.SEQ, params=cp, num_calls=num_calls)
        assert np.array_equal(c, d[f"clustering_s{seed}"]), f"clustering differs (seed {seed})"
        p, bw = B.oracle_lp_refine(g, seed, k, mbw, d[f"part_in_s{seed}"], schedule=B.SEQ, params=rp)
        assert np.array_equal(p, d[f"part_out_s{seed}"]), f"partition differs (seed {seed})"
        assert np.array_equal(bw, d[f"bw_out_s{seed}"])
        assert B.oracle_edge_cut(g, p) == int(d[f"cut_s{seed}"][0])


@pytest.mark.parametrize("name,case", [("rgg2d", "rgg2d_k4"), ("walshaw_data", "walshaw_k16"),
                                       ("rgg16_vwgt_adjwgt", "rgg16_w")])
def test_rearrange_by_degree_buckets_matches_reference(name, case):
    g0 = H.load_graph(name)
    gs, d = H.load_case(case)
    go, o2n = B.oracle_rearrange(g0)
    assert np.array_equal(go.xadj, gs.xadj) and np.array_equal(go.adjncy, gs.adjncy)
    assert np.array_equal(go.buckets, gs.buckets)
    assert np.array_equal(o2n, d["old_to_new"])
    if gs.adjwgt is not None:
        assert np.array_equal(go.adjwgt, gs.adjwgt) and np.array_equal(go.vwgt, gs.vwgt)


def test_rgg2d_shape():  # test_pykaminpar.py:78-92
    g = H.load_graph("rgg2d")
    assert g.n == 1024 and g.m == 8226 and g.vwgt is None and g.adjwgt is None


def test_same_seed_same_result_different_seed_differs():  # shm_endtoend_test.cc:189-247
    g, d = H.load_case("walshaw_k16")
    mcw = int(d["max_cluster_weight"][0])
    a = B.oracle_lp_cluster(g, 0, mcw)
    b = B.oracle_lp_cluster(g, 0, mcw)
    c = B.oracle_lp_cluster(g, 1, mcw)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_empty_and_trivial_graphs():  # shm_endtoend_test.cc:28-140
    for sched in (B.SEQ, B.SYNC):
        g = H.empty_graph(0)
        assert len(B.oracle_lp_cluster(g, 0, 10, schedule=sched)) == 0
        g = H.empty_graph(5)  # isolated nodes only
        c = B.oracle_lp_cluster(g, 0, 10, schedule=sched)
        assert len(c) == 5 and (c < 5).all()
        p, bw = B.oracle_lp_refine(g, 0, 2, [3, 3], np.array([0, 1, 0, 1, 0], np.uint32), schedule=sched)
        assert list(p) == [0, 1, 0, 1, 0] and list(bw) == [3, 2]


@pytest.mark.parametrize("sched", [B.SEQ, B.SYNC])
@pytest.mark.parametrize("maker", [lambda: H.path_graph(64), lambda: H.star_graph(40), lambda: H.complete_graph(12),
                                   lambda: H.complete_bipartite(6, 9), lambda: H.grid2d(9, 7),
                                   lambda: H.matching_graph(20)])
def test_invariants_on_fixture_families(sched, maker):
    """KASSERT-style invariants: valid ids, weight limits, consistent block weights."""
    g = maker()
    mcw = 4
    c = B.oracle_lp_cluster(g, 0, mcw, schedule=sched)
    assert (c < g.n).all() and H.cluster_weights_ok(g, c, mcw)
    k = 3
    part = (np.arange(g.n) % k).astype(np.uint32)
    mbw = np.full(k, int(1.1 * np.ceil(g.n / k)) + 1, np.int32)
    p, bw = B.oracle_lp_refine(g, 0, k, mbw, part, schedule=sched)
    assert (p < k).all() and np.array_equal(H.block_weights(g, p, k), bw) and (bw <= mbw).all()
    assert B.oracle_edge_cut(g, p) <= B.oracle_edge_cut(g, part)
