"""The C++ adapters (include/kaminpar_b200_adapters.hpp) are valid C++20, link against the C-ABI library,
fail loudly without a GPU (CPU test) and run the clusterer / contraction / refiner round trip on one
(GPU test)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "kaminpar_b200", "csrc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")


def build(tmp_path):
    exe = str(tmp_path / "adapter_smoke")
    cmd = [CXX, "-std=c++20", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cc"), "-o", exe, "-L" + LIBDIR, "-lkaminpar_b200",
           "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapter_compiles_links_and_has_no_fallback(tmp_path):
    import torch

    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_adapter_round_trip_on_gpu(tmp_path):
    """The C++ adapter path (what the reference's factories would call) gives the oracle's sync results bit for bit."""
    import numpy as np

    from kaminpar_b200.graph import CSRGraph
    from oracle import bindings as B

    exe = build(tmp_path)
    dump = str(tmp_path / "dump.txt")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, ADAPTER_DUMP=dump))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter ok" in r.stdout
    lines = open(dump).read().strip().split("\n")
    n, m = [int(x) for x in lines[0].split()]
    xadj = np.array(lines[1].split(), np.uint32)
    adj = np.array(lines[2].split(), np.uint32)
    clustering = np.array(lines[3].split(), np.uint32)
    part = np.array(lines[4].split(), np.uint32)
    bw = np.array(lines[5].split(), np.int32)
    g = CSRGraph(xadj, adj)
    assert g.n == n and g.m == m
    assert np.array_equal(clustering, B.oracle_lp_cluster(g, 0, 4, schedule=B.SYNC))       # seed 0, max weight 4
    rp = B.oracle_params(B.default_refine_params(), commit_passes=4)
    ep, ebw = B.oracle_lp_refine(g, 0, 2, np.array([9, 9], np.int32), (np.arange(n) % 2).astype(np.uint32),
                                 schedule=B.SYNC, params=rp)
    assert np.array_equal(part, ep) and np.array_equal(bw, ebw)
