"""bench.py's CPU-baseline worker (oracle/cpu_baseline_worker.py) runs in a child process and must be
stable: the multi-thread reference run (OpenMP stand-in for oneTBB) used to crash sporadically because
of a race in the stand-in's enumerable_thread_specific (oracle/ref_shim/tbb/shim_core.h)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from kaminpar_b200.graph import rmat
from oracle import bindings as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_worker(path, mode):
    r = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline_worker", path, mode, "1", "0"], cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


def test_worker_runs_lp_and_contraction(tmp_path):
    g = B.oracle_rearrange(rmat(14, 16, 3))[0]
    path = str(tmp_path / "g.npz")
    np.savez(path, xadj=g.xadj, adjncy=g.adjncy, k=np.array([8]))
    for attempt in range(4):  # the old race hit roughly every second run
        d = run_worker(path, "lp")
        assert d["value"] > 0 and d["units"] > 0
        assert d["kind"] == ("reference" if B.have_reference() else "port")
        if B.have_parallel_reference() and (os.cpu_count() or 1) > 1:
            assert d["cores"] > 1
    d = run_worker(path, "contraction")
    assert d["units"] == g.m and d["cores"] == 1
