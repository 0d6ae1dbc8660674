"""CPU-baseline worker (TEST / BENCH INFRASTRUCTURE ONLY; executed by bench.py's cpu_baseline and
`--impl reference` legs in a child process).

Why a child process: isolation. The reference runs on the OpenMP stand-in for oneTBB (system libgomp),
bench.py's process has torch loaded with its own OpenMP runtime and busy worker threads; a crash of
the baseline must not take the GPU measurement down (bench.py falls back to the serial stand-in and
says so in `cores`). The flaky SIGSEGV seen on the GPU box was a race in OUR stand-in (the
enumerable_thread_specific slot table was allocated lazily while a whole team called local() at
once), fixed in oracle/ref_shim/tbb/shim_core.h; OMP_STACKSIZE is raised as a precaution because the
reference keeps sizeable per-thread buffers on the stack.

    python -m oracle.cpu_baseline_worker <graph.npz> <lp|contraction> <steps> <warmup> [serial]

graph.npz: xadj, adjncy (uint32), k. Prints one JSON line:
{"value": units/s, "seconds_per_step": s, "kind": "reference"|"port", "cores": c, "units": u, "desc": "..."}
"""
import json
import os
import sys
import time

os.environ.setdefault("OMP_STACKSIZE", "64M")  # before anything loads libgomp
os.environ.pop("OMP_NUM_THREADS", None)  # torchrun sets it to 1; the thread count is chosen below

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kaminpar_b200.graph import CSRGraph  # noqa: E402  (numpy only)
from oracle import bindings as B  # noqa: E402


SWEEP = []  # thread counts tried for the reference and their time per step


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_lp(g, k, steps, warmup, serial):
    mcw = B.oracle_max_cluster_weight(g, k)
    _, st = B.oracle_lp_cluster(g, 0, mcw, schedule=B.SEQ, return_stats=True)
    edges = int(st[0].edges_scanned)  # scanned-edge count of the sequential schedule
    cores = 1
    if B.have_parallel_reference() and not serial:
        kind, fn = "reference", (lambda: B.ref_lp_cluster(g, 0, mcw, parallel=True))
        # the thread count that is fastest for the reference on this box (all cores is not always best:
        # the graph is first-touched by one thread), so that the baseline is not handicapped
        # cores this process may run on -- NOT OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1 to
        # its children, which made the round-1 reference arm single-threaded at N > 1
        avail = host_cores()
        best = None
        # The stand-in's parallel_for (an OpenMP dynamic loop over the reference's own chunk claims) stops scaling
        # early: R-MAT 22 on a 128-core host measured 1.5-1.8 s/step at 16 threads, 2.2 s at 32, 3.7-4.8 s at 64 and
        # 150-160 s at 128 (gpurun_out/bench_ref.json, round 2) -- so the sweep stops at 64 threads and reports them all.
        for t in sorted({min(avail, 64), min(avail, 32), min(avail, 16), min(avail, 8)}, reverse=True):
            B.ref_omp().kmpref_set_num_threads(t)
            fn()
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            SWEEP.append({"threads": t, "seconds_per_step": dt})
            if best is None or dt < best[0]:
                best = (dt, t)
        cores = best[1]
        B.ref_omp().kmpref_set_num_threads(cores)
    elif B.have_reference():
        kind, fn = "reference", (lambda: B.ref_lp_cluster(g, 0, mcw))
    else:
        kind, fn = "port", (lambda: B.oracle_lp_cluster(g, 0, mcw, schedule=B.SEQ))
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    desc = (f"n={g.n} m={g.m} k={k}, LPClustering.compute_clustering on {cores} thread(s) "
            f"({edges} scanned edges/step counted by the 1-thread schedule)")
    return edges, dt, kind, cores, desc


def run_contraction(g, k, steps, warmup):
    from oracle import contraction_oracle as CO

    cl = B.oracle_lp_cluster(g, 0, B.oracle_max_cluster_weight(g, k), schedule=B.SYNC)
    if B.have_reference():
        kind, fn = "reference", (lambda: B.ref_contract(g, cl, 1))
    else:
        kind, fn = "port", (lambda: CO.contract(g.xadj, g.adjncy, None, None, cl))
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return g.m, dt, kind, 1, f"n={g.n} m={g.m}, contract_clustering (UNBUFFERED) of the LP clustering on 1 thread(s)"


def main():
    path, mode, steps, warmup = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    serial = len(sys.argv) > 5 and sys.argv[5] == "serial"
    d = np.load(path)
    g = CSRGraph(d["xadj"], d["adjncy"], sorted=True)
    k = int(d["k"][0])
    if mode == "lp":
        units, dt, kind, cores, desc = run_lp(g, k, steps, warmup, serial)
    else:
        units, dt, kind, cores, desc = run_contraction(g, k, steps, warmup)
    print(json.dumps({"value": units / dt, "seconds_per_step": dt, "kind": kind, "cores": cores, "units": units,
                      "desc": desc, "host_cores": host_cores(), "thread_sweep": SWEEP}), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
