"""Generate the golden fixtures under tests/golden/ (run in the authoring container only).

Inputs come from the reference checkout's data files (graphs, not source code); expected outputs
come from the UNMODIFIED reference compiled against the serial oneTBB stand-in
(oracle/_ref/libkaminpar_ref.so, `make -C oracle ref`), i.e. the reference's own LP clusterer and
LP refiner at one thread. The oracle's `seq` schedule must reproduce every vector bit for bit
(tests/test_oracle_golden.py); that is what pins the oracle.

    python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from kaminpar_b200.graph import CSRGraph, grid3d, random_weights, read_metis, rgg2d, rmat, road_like  # noqa: E402
from oracle import bindings as B  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def walshaw_data() -> CSRGraph:
    xs = open(f"{REF}/tests/endtoend/data.graph.xadj").read()
    ad = open(f"{REF}/tests/endtoend/data.graph.adjncy").read()
    xadj = np.array([int(t) for t in re.findall(r"\d+", xs)], np.uint32)
    adj = np.array([int(t) for t in re.findall(r"\d+", ad)], np.uint32)
    return CSRGraph(xadj, adj)


def star(n: int) -> CSRGraph:
    xadj = np.zeros(n + 1, np.int64)
    xadj[1] = n - 1
    xadj[2:] = n - 1 + np.arange(1, n)
    adj = np.concatenate([np.arange(1, n), np.zeros(n - 1, np.int64)])
    return CSRGraph(xadj.astype(np.uint32), adj.astype(np.uint32))


def save_graph(name: str, g: CSRGraph):
    d = {"xadj": g.xadj, "adjncy": g.adjncy}
    if g.vwgt is not None:
        d["vwgt"] = g.vwgt
    if g.adjwgt is not None:
        d["adjwgt"] = g.adjwgt
    np.savez_compressed(os.path.join(OUT, f"graph_{name}.npz"), **d)


def main():
    assert B.have_reference(), "build oracle/_ref first: make -C oracle ref"
    inputs = {
        "rgg2d": read_metis(f"{REF}/misc/rgg2d.metis"),                   # BASELINE config 1
        "rgg16": read_metis(f"{REF}/tests/io/rgg16.metis"),
        "rgg16_vwgt_adjwgt": read_metis(f"{REF}/tests/io/rgg16-vwgt-adjwgt.metis"),
        "walshaw_data": walshaw_data(),                                   # shm_endtoend_test.cc:18-24
    }
    for name, g in inputs.items():
        save_graph(name, g)

    cases = []  # (case name, graph, sorted?, k, cluster params, refine params, seeds, num_calls)
    cp_geo, rp_geo = B.default_cluster_params(), B.default_refine_params()
    cp_geo.tie_breaking = 0
    rp_geo.tie_breaking = 0
    cases.append(("rgg2d_k4", inputs["rgg2d"], True, 4, None, None, (0, 1), 1))
    cases.append(("rgg2d_k2", inputs["rgg2d"], True, 2, None, None, (0,), 1))
    cases.append(("rgg16_w", inputs["rgg16_vwgt_adjwgt"], True, 2, None, None, (0,), 1))
    cases.append(("walshaw_k16", inputs["walshaw_data"], True, 16, None, None, (0, 3), 1))
    cases.append(("walshaw_unsorted", inputs["walshaw_data"], False, 16, None, None, (0,), 1))
    cases.append(("walshaw_geometric", inputs["walshaw_data"], True, 8, cp_geo, rp_geo, (0,), 1))
    cases.append(("walshaw_3calls", inputs["walshaw_data"], True, 16, None, None, (5,), 3))
    cases.append(("grid12", grid3d(12), True, 8, None, None, (0,), 1))
    cases.append(("rmat13_w", random_weights(rmat(13, 8, 2), 5, max_vwgt=3, max_adjwgt=5), False, 8, None, None, (0,), 1))
    cases.append(("rmat14", rmat(14, 16, 1), True, 16, None, None, (0,), 1))
    cases.append(("road60", road_like(60, 2, 0.3, 0.5), True, 4, None, None, (0,), 1))
    cases.append(("star30000", star(30000), True, 2, None, None, (0,), 1))  # second phase + two-hop
    for ths in (1, 3, 4):
        cp = B.default_cluster_params()
        cp.two_hop_strategy = ths
        cp.isolated_nodes_strategy = 1
        cases.append((f"star4000_twohop{ths}", star(4000), False, 2, cp, None, (1,), 1))

    for name, g0, sort, k, cp, rp, seeds, num_calls in cases:
        if sort:
            g, o2n = B.ref_rearrange(g0)
        else:
            g, o2n = g0, np.arange(g0.n, dtype=np.uint32)
        out = {"xadj": g.xadj, "adjncy": g.adjncy, "sorted": np.array([1 if g.sorted else 0]),
               "k": np.array([k]), "seeds": np.array(seeds), "old_to_new": o2n,
               "num_calls": np.array([num_calls])}
        if g.buckets is not None:
            out["buckets"] = g.buckets
        if g.vwgt is not None:
            out["vwgt"] = g.vwgt
        if g.adjwgt is not None:
            out["adjwgt"] = g.adjwgt
        mcw = B.ref_max_cluster_weight(g, k)
        mbw = B.ref_max_block_weights(g, k)
        out["max_cluster_weight"] = np.array([mcw])
        out["max_block_weights"] = mbw
        for p, tag in ((cp, "cparams"), (rp, "rparams")):
            if p is not None:
                out[tag] = np.array([p.num_iterations, p.impl, p.tie_breaking, p.two_hop_strategy,
                                     p.isolated_nodes_strategy])
        for seed in seeds:
            out[f"clustering_s{seed}"] = B.ref_lp_cluster(g, seed, mcw, params=cp, num_calls=num_calls)
            rng = np.random.default_rng(seed)
            part = rng.integers(0, k, g.n).astype(np.uint32)
            p2, bw = B.ref_lp_refine(g, seed, k, mbw, part, params=rp)
            out[f"part_in_s{seed}"] = part
            out[f"part_out_s{seed}"] = p2
            out[f"bw_out_s{seed}"] = bw
            out[f"cut_s{seed}"] = np.array([B.ref_edge_cut(g, k, p2)])
        np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
        print("wrote", name, g.n, g.m)


if __name__ == "__main__":
    main()
