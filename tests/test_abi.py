"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/kaminpar_b200_lp.h declares, and fails loudly (no fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

from kaminpar_b200 import lp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for header in ("kaminpar_b200_lp.h", "kaminpar_b200_contraction.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b(kmp_[a-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    lib = lp.load_library()
    syms = declared_symbols()
    assert len(syms) >= 35 and "kmp_contract_clustering" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert lib.kmp_lp_abi_version() == 3


def test_config_struct_layout_matches_header():
    cfg = lp.KmpConfig()
    lp.load_library().kmp_lp_default_config(0, ctypes.byref(cfg))
    assert (cfg.num_iterations, cfg.large_degree_threshold, cfg.max_num_neighbors) == (5, 0xFFFFFFFF, 0xFFFFFFFF)
    assert cfg.two_hop_strategy == 2 and abs(cfg.two_hop_threshold - 0.5) < 1e-12
    assert cfg.isolated_nodes_strategy == 3 and cfg.sync_subrounds == 8 and cfg.sync_granule_log2 == 4
    assert cfg.sync_commit_passes == 1 and cfg.device == -1
    lp.load_library().kmp_lp_default_config(1, ctypes.byref(cfg))
    assert cfg.impl == 0 and cfg.sync_commit_passes == 4


def test_no_silent_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        lp.LPClustering(lp.CoarseningContext())


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under kaminpar_b200/ may import, link or load it."""
    import re

    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|liblp_oracle|libkaminpar_ref|oracle/bindings|#include\s+\"[^\"]*oracle",
                     re.M)
    for root, _, files in os.walk(os.path.join(ROOT, "kaminpar_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(root, f)).read()
                assert not pat.search(text), f


def test_host_mirror_scalars():
    """PartitionContext.setup / compute_max_cluster_weight mirror vs values from the reference
    (stored in the golden files)."""
    import numpy as np

    from tests import helpers as H

    for name in ("rgg2d_k4", "walshaw_k16", "rmat13_w", "rmat14"):
        g, d = H.load_case(name)
        k = int(d["k"][0])
        ctx = lp.create_default_context()
        ctx.partition.setup(g, k, 0.03)
        assert np.array_equal(ctx.partition.max_block_weights(), d["max_block_weights"])
        mcw = lp.compute_max_cluster_weight(ctx.coarsening, ctx.partition, g.n, g.total_node_weight())
        assert mcw == int(d["max_cluster_weight"][0])
