"""CPU check of the register sorting network the thread-per-vertex sweep kernels run (kaminpar_b200/csrc/lp_sortnet.cuh
compiled with g++ by tests/cpp/sortnet_host_check.cc -- test infrastructure; the product only runs it inside
sweep_thread<N>): 0-1 principle (exhaustive for N = 8 / 16), weights travel with their keys, Batcher's
compare-exchange counts. The -m gpu parity tests run the same source on the device."""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "sortnet_host_check.cc")
LIB = os.path.join(HERE, "cpp", "libsortnet_host_check.so")
HDR = os.path.join(os.path.dirname(HERE), "kaminpar_b200", "csrc", "lp_sortnet.cuh")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


@pytest.mark.parametrize("n,pairs", [(8, 19), (16, 63), (32, 191), (64, 543)])
def test_network_is_batchers(lib, n, pairs):
    assert lib.sortnet_pairs(n) == pairs


@pytest.mark.parametrize("n", [8, 16, 32, 64])
def test_zero_one_principle(lib, n):
    assert lib.sortnet_zero_one(n, 12345) == 0


@pytest.mark.parametrize("n", [8, 16, 32, 64])
def test_weights_travel_with_their_keys(lib, n):
    assert lib.sortnet_weighted(n, 7, 20000) == 0
