"""-m gpu: schedule KMP_SCHEDULE_SEQ_STRICT through the C ABI == the UNMODIFIED reference, bit for bit.

tests/golden/ref_*.npz hold what the reference's LPClustering / LabelPropagationRefiner (one thread, real
libstdc++ random facilities) computed for these inputs (tests/golden/make_golden.py). The device engine
(kaminpar_b200/csrc/lp_strict.cuh) walks the same chunk order with the same draws, so the label vector,
the block weights and the cut must be identical -- this is BASELINE.json's "same partition vector" for the
configuration in which the reference itself is deterministic (config 1, misc/rgg2d.metis k=4, and 14 more).
"""
import numpy as np
import pytest

from kaminpar_b200 import lp
from oracle import bindings as B
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _ctx(d, seed):
    ctx = lp.create_default_context()
    ctx.engine.seed = seed
    ctx.engine.schedule = "seq_strict"
    c = ctx.coarsening.clustering.lp
    if "cparams" in d:
        c.num_iterations, c.impl, c.tie_breaking_strategy, c.two_hop_strategy, c.isolated_nodes_strategy = \
            [int(x) for x in d["cparams"]]
    r = ctx.refinement.lp
    if "rparams" in d:
        r.num_iterations, r.impl, r.tie_breaking_strategy = [int(x) for x in d["rparams"][:3]]
    return ctx


@pytest.mark.parametrize("name", H.golden_cases())
def test_strict_schedule_is_bit_identical_to_the_reference(name):
    g, d = H.load_case(name)
    k = int(d["k"][0])
    mcw = int(d["max_cluster_weight"][0])
    num_calls = int(d["num_calls"][0])
    for seed in d["seeds"]:
        seed = int(seed)
        ctx = _ctx(d, seed)
        ctx.partition.setup(g, k, 0.03)
        assert np.array_equal(ctx.partition.max_block_weights(), d["max_block_weights"])
        clusterer = lp.LPClustering(ctx.coarsening, ctx.engine)
        clusterer.set_max_cluster_weight(mcw)
        exp = d[f"clustering_s{seed}"]
        for call in range(num_calls):  # same object: the random stream continues (overlay coarsener)
            c = clusterer.compute_clustering(g)
            assert np.array_equal(c, exp if num_calls == 1 else exp[call]), f"clustering differs (seed {seed}, call {call})"
        part = np.ascontiguousarray(d[f"part_in_s{seed}"], np.uint32).copy()
        p_graph = lp.PartitionedGraph(g, k, part)
        refiner = lp.LabelPropagationRefiner(ctx)
        refiner.initialize(p_graph)
        refiner.refine(p_graph, ctx.partition)
        assert np.array_equal(p_graph.partition, d[f"part_out_s{seed}"]), f"partition differs (seed {seed})"
        assert np.array_equal(p_graph.block_weights(), d[f"bw_out_s{seed}"])
        assert B.oracle_edge_cut(g, p_graph.partition) == int(d[f"cut_s{seed}"][0])


def test_strict_counts_and_limits():
    """statistics of the strict engine == the seq oracle's; inputs above the size limit are refused"""
    g, d = H.load_case("walshaw_k16")
    mcw = int(d["max_cluster_weight"][0])
    ctx = _ctx(d, 0)
    clusterer = lp.LPClustering(ctx.coarsening, ctx.engine)
    clusterer.set_max_cluster_weight(mcw)
    c = clusterer.compute_clustering(g)
    exp, st = B.oracle_lp_cluster(g, 0, mcw, schedule=B.SEQ, return_stats=True)
    assert np.array_equal(c, exp)
    gs = clusterer.last_stats
    assert gs.moved_list() == list(st[0].moved[: st[0].iterations])
    assert gs.edges_scanned == st[0].edges_scanned and gs.nodes_visited == st[0].nodes_visited
    assert gs.num_clusters == st[0].num_clusters and gs.two_hop_ran == st[0].two_hop_ran


def test_sync_schedule_refuses_order_dependent_options():
    """GEOMETRIC tie-breaking, the global two-hop variants and relabel_before_second_phase have no order-free
    restatement: KMP_ERR_UNSUPPORTED under the sync schedule (never silently mapped), accepted by seq_strict."""
    ctx = lp.create_default_context()
    for field, value in (("tie_breaking_strategy", 0), ("two_hop_strategy", 1), ("two_hop_strategy", 3),
                         ("relabel_before_second_phase", True)):
        c = lp.create_default_context().coarsening
        setattr(c.clustering.lp, field, value)
        with pytest.raises(RuntimeError, match="-4"):
            lp.LPClustering(c, ctx.engine)
        eng = lp.EngineContext(schedule="seq_strict")
        if field != "relabel_before_second_phase":
            lp.LPClustering(c, eng)
