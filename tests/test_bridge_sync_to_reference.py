"""Bridge between the `sync` selection rule (what the GPU computes, T0/T1) and the REFERENCE.

The `seq` oracle's select_best_cluster code (oracle/lp_oracle.cc ClusterPolicy / RefinePolicy) is pinned
bit-for-bit to the unmodified reference by tests/test_oracle_golden.py. Here the SAME code is run on frozen
mid-run states (lpo_seq_select_all) next to the sync rule (lpo_sync_select_all, separate code):

  * a vertex whose set of maximal feasible candidates is a singleton (or empty) has ONE answer; sync must give it;
  * otherwise the reference draws uniformly from that set (lp_clusterer.cc:238-241, lp_refiner.cc:233-236);
    sync's hash-chosen target must be a member of exactly that set.

Same for the favored cluster of the clusterer (lp_clusterer.cc:243-248). Together with T0 (GPU == sync on
frozen state) this ties the GPU's per-vertex decision to the reference's selection rule.
"""
import numpy as np
import pytest

from oracle import bindings as B
from tests.helpers import golden_cases, load_case


def cluster_states(g, mcw):
    """frozen clusterer states: initial singletons, and the sync schedule after 1, 2 and 4 rounds"""
    w = np.ones(g.n, np.int64) if g.vwgt is None else g.vwgt.astype(np.int64)
    yield "init", np.arange(g.n, dtype=np.uint32), w.astype(np.int32)
    for rounds in (1, 2, 4):
        p = B.default_cluster_params()
        p.num_iterations = rounds
        p.two_hop_strategy = 0
        p.isolated_nodes_strategy = 0
        labels = B.oracle_lp_cluster(g, 0, mcw, schedule=B.SYNC, params=p)
        cw = np.bincount(labels, weights=w, minlength=g.n).astype(np.int32)
        yield f"sync{rounds}", labels, cw


@pytest.mark.parametrize("case", golden_cases())
def test_sync_cluster_selection_is_the_references(case):
    g, d = load_case(case)
    if g.n == 0:
        pytest.skip("empty")
    k = int(d["k"][0]) if "k" in d else 8
    mcw = max(int(B.oracle_max_cluster_weight(g, k)), 1)
    checked = singles = 0
    for name, labels, cw in cluster_states(g, mcw):
        for it in (0, 3):
            tgt, fav = B.oracle_sync_select_all(0, g, labels, cw, max_cluster_weight=mcw, seed=1, call=0, iteration=it)
            fav_chk = np.where(fav == 0xFFFFFFFF, labels, fav).astype(np.uint32)
            r = B.oracle_seq_select_all(0, g, labels, cw, max_cluster_weight=mcw, check_target=tgt,
                                        check_favored=fav_chk)
            bad = np.nonzero(r["check_in_ties"] == 0)[0]
            assert bad.size == 0, (case, name, it, bad[:5], tgt[bad[:5]], r["target"][bad[:5]])
            single = r["num_ties"] <= 1
            assert np.array_equal(tgt[single], r["target"][single]), (case, name, it)
            # favored: stored for the same vertices, member of the reference's tie set, equal when unique
            assert np.array_equal(fav == 0xFFFFFFFF, r["favored"] == 0xFFFFFFFF), (case, name, it)
            assert (r["check_fav_in_ties"] == 1).all(), (case, name, it)
            fs = (r["num_fav_ties"] <= 1) & (fav != 0xFFFFFFFF)
            assert np.array_equal(fav[fs], r["favored"][fs]), (case, name, it)
            checked += g.n
            singles += int(single.sum())
    assert checked > 0 and singles > 0


@pytest.mark.parametrize("case", golden_cases())
def test_sync_refine_selection_is_the_references(case):
    g, d = load_case(case)
    if g.n == 0:
        pytest.skip("empty")
    w = np.ones(g.n, np.int64) if g.vwgt is None else g.vwgt.astype(np.int64)
    for k in (2, 4, 16):
        if k > g.n:
            continue
        rng = np.random.default_rng(k)
        part = rng.integers(0, k, g.n).astype(np.uint32)
        mbw = B.oracle_max_block_weights(g, k)
        states = [("random", part)]
        p = B.default_refine_params()
        p.num_iterations = 2
        part2, _ = B.oracle_lp_refine(g, 0, k, mbw, part, schedule=B.SYNC, params=p)
        states.append(("sync2", part2))
        for name, lab in states:
            bw = np.bincount(lab, weights=w, minlength=k).astype(np.int32)
            for min_w in (None, np.full(k, int(bw.min()), np.int32)):
                tgt, _ = B.oracle_sync_select_all(1, g, lab, bw, max_weights=mbw, min_weights=min_w, seed=2, iteration=1)
                r = B.oracle_seq_select_all(1, g, lab, bw, max_weights=mbw, min_weights=min_w, check_target=tgt)
                bad = np.nonzero(r["check_in_ties"] == 0)[0]
                assert bad.size == 0, (case, k, name, bad[:5], tgt[bad[:5]], r["target"][bad[:5]])
                single = r["num_ties"] <= 1
                assert np.array_equal(tgt[single], r["target"][single]), (case, k, name)
