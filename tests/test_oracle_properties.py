"""CPU property tests of the oracle's two schedules on random multigraphs (hypothesis): the
executable invariants inside the reference's hot path (valid ids, block weights equal the
sum of node weights, limits respected, cut never increases for the refiner's positive-gain moves
in the sequential schedule)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import bindings as B
from tests import helpers as H


@st.composite
def graphs(draw):
    n = draw(st.integers(1, 50))
    m_und = draw(st.integers(0, 150))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    src, dst = rng.integers(0, n, m_und), rng.integers(0, n, m_und)
    vw = rng.integers(1, 5, n) if draw(st.booleans()) else None
    ew = rng.integers(1, 6, m_und).tolist() if draw(st.booleans()) else None
    return H.from_edges(n, list(zip(src.tolist(), dst.tolist())), vwgt=vw, ew=ew)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(g=graphs(), seed=st.integers(0, 100), mcw=st.integers(1, 30), sched=st.sampled_from([B.SEQ, B.SYNC]))
def test_clustering_invariants(g, seed, mcw, sched):
    c = B.oracle_lp_cluster(g, seed, mcw, schedule=sched)
    assert len(c) == g.n and (c < max(g.n, 1)).all()
    assert H.cluster_weights_ok(g, c, mcw)
    assert np.array_equal(c, B.oracle_lp_cluster(g, seed, mcw, schedule=sched))  # deterministic


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(g=graphs(), seed=st.integers(0, 100), k=st.integers(2, 7), sched=st.sampled_from([B.SEQ, B.SYNC]))
def test_refinement_invariants(g, seed, k, sched):
    rng = np.random.default_rng(seed)
    part = rng.integers(0, k, g.n).astype(np.uint32)
    bw0 = H.block_weights(g, part, k)
    mbw = np.full(k, int(bw0.max() * 1.2) + 1, np.int32)
    rp = B.oracle_params(B.default_refine_params(), commit_passes=4)
    p, bw = B.oracle_lp_refine(g, seed, k, mbw, part, schedule=sched, params=rp)
    assert (p < k).all() and np.array_equal(H.block_weights(g, p, k), bw) and (bw <= mbw).all()
    if sched == B.SEQ:  # sequential moves have non-negative gain each
        assert B.oracle_edge_cut(g, p) <= B.oracle_edge_cut(g, part)


@pytest.mark.parametrize("name", ["rgg2d_k4", "walshaw_k16", "rmat13_w"])
def test_sync_refine_with_min_weights_never_overshoots_max(name):
    """arrivals accepted on the credit of departures that the min-weight thinning later revokes
    must not push a block above its maximum (the reference's move_block_weight checks both limits atomically,
    partitioned_graph.h:397-428)."""
    from tests.helpers import load_case

    g, _ = load_case(name)
    w = np.ones(g.n, np.int64) if g.vwgt is None else g.vwgt.astype(np.int64)
    for k in (2, 8):
        part = (np.arange(g.n) * k // g.n).astype(np.uint32)
        bw0 = np.bincount(part, weights=w, minlength=k).astype(np.int64)
        mbw = np.full(k, int(1.01 * bw0.max()) + int(w.max()), np.int32)
        minw = (0.98 * bw0).astype(np.int32)
        for passes in (1, 4):
            rp = B.oracle_params(B.default_refine_params(), commit_passes=passes)
            ep, ebw = B.oracle_lp_refine(g, 5, k, mbw, part, schedule=B.SYNC, params=rp, min_block_weights=minw)
            assert np.array_equal(np.bincount(ep, weights=w, minlength=k).astype(np.int64), ebw.astype(np.int64))
            assert (ebw >= minw).all() and (ebw <= np.maximum(mbw, bw0)).all(), (name, k, passes, ebw, mbw, minw)
