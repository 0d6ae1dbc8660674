"""CPU property tests of the oracle's two schedules on random multigraphs (hypothesis): the
executable invariants inside the reference's hot path (valid ids, block weights equal the
sum of node weights, limits respected, cut never increases for the refiner's positive-gain moves
in the sequential schedule)."""
import numpy as np
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
