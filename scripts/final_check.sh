#!/bin/bash
# last call of the round: GPU suite, smoke(), bench table, default bench line (no profiler)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/round2_table.sh 2>&1 | tail -9
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c1-300 gpurun_out/r2_bench_default.json
