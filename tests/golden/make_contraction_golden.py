"""Golden fixtures of the contraction row (run in the authoring container only).

For the graphs and reference LP clusterings already stored in tests/golden/ref_<case>.npz, store what the
UNMODIFIED reference's contract_clustering returns (default algorithm UNBUFFERED; raw, i.e. with the
reference's own coarse numbering and adjacency order) -> tests/golden/contract_<case>.npz. The oracle
(oracle/contraction_oracle.py) must reproduce every one of them after canonicalize()
(tests/test_contraction_oracle.py); that is what pins it.

    python tests/golden/make_contraction_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bindings as B  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = ["rgg2d_k4", "rgg16_w", "walshaw_k16", "walshaw_unsorted", "rmat13_w", "grid12", "road60", "star30000"]


def main():
    assert B.have_reference(), "build oracle/_ref first: make -C oracle ref"
    for name in CASES:
        g, d = H.load_case(name)
        seed = int(d["seeds"][0])
        cl = d[f"clustering_s{seed}"]
        cl = cl if cl.ndim == 1 else cl[0]
        r = B.ref_contract(g, cl, 1)
        np.savez_compressed(os.path.join(OUT, f"contract_{name}.npz"), clustering=cl.astype(np.uint32),
                            c_n=np.array([r["c_n"]]), c_xadj=r["c_xadj"], c_adjncy=r["c_adjncy"], c_vwgt=r["c_vwgt"],
                            c_adjwgt=r["c_adjwgt"], mapping=r["mapping"])
        print("wrote", name, g.n, g.m, "->", r["c_n"], len(r["c_adjncy"]))


if __name__ == "__main__":
    main()
