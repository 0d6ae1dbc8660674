"""CPU checks of the KMP_SCHEDULE_SEQ_STRICT engine's SOURCE (kaminpar_b200/csrc/lp_strict.cuh compiled with g++
by tests/cpp/strict_host_check.cc -- test infrastructure, the product only runs it as a CUDA kernel):

  * its restatements of libstdc++'s std::mt19937 / std::uniform_int_distribution / std::shuffle produce the
    same draws as the real facilities (the reference's Random wraps exactly these, kaminpar-common/random.h:64-88);
  * the sequential engine reproduces the UNMODIFIED reference's outputs (tests/golden/ref_*.npz) bit for bit.

The -m gpu twin (tests/test_gpu_strict.py) runs the same engine on the device through the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "strict_host_check.cc")
LIB = os.path.join(HERE, "cpp", "libstrict_host_check.so")
HDR = os.path.join(os.path.dirname(HERE), "kaminpar_b200", "csrc", "lp_strict.cuh")


class StrictStats(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("moved", C.c_uint32 * 64), ("edges_scanned", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("num_clusters", C.c_uint32), ("two_hop_ran", C.c_uint32)]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def test_rng_restatement_matches_libstdcxx(lib):
    for seed in (0, 1, 42, -7, 2**31 - 1):
        assert lib.strict_check_rng(C.c_int(seed), C.c_int(40)) == 0
        assert lib.strict_check_rng_init(C.c_int(seed)) == 0


def _params(d, tag, default):
    if tag in d:
        return [int(x) for x in d[tag]]  # it, impl, tie, two_hop, iso
    return default


@pytest.mark.parametrize("name", H.golden_cases())
def test_strict_engine_reproduces_reference(lib, name):
    g, d = H.load_case(name)
    k = int(d["k"][0])
    mcw = int(d["max_cluster_weight"][0])
    mbw = np.ascontiguousarray(d["max_block_weights"], np.int32)
    num_calls = int(d["num_calls"][0])
    it, impl, tie, ths, iso = _params(d, "cparams", [5, 1, 1, 2, 3])
    rit, rimpl, rtie, _, _ = _params(d, "rparams", [5, 0, 1, 0, 0])
    for seed in d["seeds"]:
        seed = int(seed)
        out = np.zeros(g.n * num_calls, np.uint32)
        stats = (StrictStats * num_calls)()
        lib.strict_host_cluster(C.c_uint32(g.n), C.c_uint32(g.m), _p(g.xadj), _p(g.adjncy), _p(g.vwgt), _p(g.adjwgt),
                                C.c_int(1 if g.sorted else 0), C.c_int(seed), C.c_int32(mcw), C.c_uint32(0), None,
                                C.c_uint32(it), C.c_uint32(0xFFFFFFFF), C.c_uint32(0xFFFFFFFF), C.c_int(impl),
                                C.c_int(tie), C.c_int(ths), C.c_double(0.5), C.c_int(iso), C.c_int(num_calls),
                                _p(out), stats)
        exp = d[f"clustering_s{seed}"]
        got = out if num_calls == 1 else out.reshape(num_calls, g.n)
        assert np.array_equal(got, exp), f"clustering differs (seed {seed})"
        part = np.ascontiguousarray(d[f"part_in_s{seed}"], np.uint32).copy()
        bw = np.zeros(k, np.int32)
        rst = StrictStats()
        lib.strict_host_refine(C.c_uint32(g.n), C.c_uint32(g.m), _p(g.xadj), _p(g.adjncy), _p(g.vwgt), _p(g.adjwgt),
                               C.c_int(1 if g.sorted else 0), C.c_int(seed), C.c_uint32(k), _p(mbw), None, None,
                               C.c_uint32(rit), C.c_uint32(0xFFFFFFFF), C.c_uint32(0xFFFFFFFF), C.c_int(rimpl),
                               C.c_int(rtie), _p(part), _p(bw), C.byref(rst))
        assert np.array_equal(part, d[f"part_out_s{seed}"]), f"partition differs (seed {seed})"
        assert np.array_equal(bw, d[f"bw_out_s{seed}"])
