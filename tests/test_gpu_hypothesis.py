"""Property-based parity: random small multigraphs (parallel edges, self-loops, isolated vertices,
random node / edge weights, random k and block-weight limits) -- GPU == oracle `sync`, plus the
reference's KASSERT-style invariants (valid ids, weight limits, consistent block weights)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from kaminpar_b200 import lp
from kaminpar_b200.graph import CSRGraph
from oracle import bindings as B
from tests import helpers as H

pytestmark = pytest.mark.gpu


@st.composite
def graphs(draw):
    n = draw(st.integers(2, 60))
    m_und = draw(st.integers(0, 200))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    hub = draw(st.booleans())
    src = rng.integers(0, n, m_und)
    dst = rng.integers(0, n, m_und)
    if hub and n > 4:  # a vertex adjacent to (almost) everything, several times
        extra = np.arange(1, n)
        src = np.concatenate([src, np.zeros(len(extra) * 3, np.int64)])
        dst = np.concatenate([dst, np.tile(extra, 3)])
    weighted_e = draw(st.booleans())
    weighted_v = draw(st.booleans())
    ew = rng.integers(1, 6, len(src)) if weighted_e else None
    edges = list(zip(src.tolist(), dst.tolist()))
    g = H.from_edges(n, edges, vwgt=rng.integers(1, 5, n) if weighted_v else None, ew=None if ew is None else ew.tolist())
    return g


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(g=graphs(), seed=st.integers(0, 1000), mcw=st.integers(1, 40))
def test_cluster_matches_oracle_on_random_multigraphs(g, seed, mcw):
    ctx = lp.create_default_context()
    ctx.engine.seed = seed
    clusterer = lp.LPClustering(ctx.coarsening, ctx.engine)
    clusterer.set_max_cluster_weight(mcw)
    c = clusterer.compute_clustering(g)
    expect = B.oracle_lp_cluster(g, seed, mcw, schedule=B.SYNC)
    assert np.array_equal(c, expect)
    assert (c < g.n).all() and H.cluster_weights_ok(g, c, mcw)


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(g=graphs(), seed=st.integers(0, 1000), k=st.integers(2, 9), slack=st.floats(0.0, 0.5))
def test_refine_matches_oracle_on_random_multigraphs(g, seed, k, slack):
    rng = np.random.default_rng(seed)
    part = rng.integers(0, k, g.n).astype(np.uint32)
    bw0 = H.block_weights(g, part, k)
    mbw = np.maximum(1, (bw0.max() * (1.0 + slack)).astype(np.int64) * np.ones(k, np.int64)).astype(np.int32)
    ctx = lp.create_default_context()
    ctx.engine.seed = seed
    ctx.partition.setup(g, [int(x) for x in mbw])
    p_graph = lp.PartitionedGraph(g, k, part)
    refiner = lp.LabelPropagationRefiner(ctx)
    refiner.initialize(p_graph)
    refiner.refine(p_graph, ctx.partition)
    rp = B.oracle_params(B.default_refine_params(), commit_passes=4)
    ep, ebw = B.oracle_lp_refine(g, seed, k, mbw, part, schedule=B.SYNC, params=rp)
    assert np.array_equal(p_graph.partition, ep) and np.array_equal(p_graph.block_weights(), ebw)
    assert np.array_equal(H.block_weights(g, ep, k), ebw)
    # blocks that started within their limit stay within it (moves never overshoot)
    assert ((ebw <= mbw) | (bw0 > mbw)).all()
