#!/bin/bash
# Round-2 final evidence on ONE GPU (under gpurun): GPU test suite, bench table, default bench line, ncu launch
# list of the default command, --set full capture of the hub kernels on R-MAT 24 (CSV pages only: gpurun_out/ is
# capped at 64 MiB). Numbers printed by bench.py under ncu are not bench values.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
bash scripts/round2_table.sh 2>&1 | tail -9
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c1-400 gpurun_out/r2_bench_default.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|commit_|k_[a-z]|reset_" -c 4000 --csv \
  --log-file gpurun_out/r2_launches_rmat22.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_a.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"sweep_hub" -s 6 -c 6 \
  -o /tmp/r2_hub_rmat24 python bench.py --workload rmat24 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_c.log 2>&1
ncu -i /tmp/r2_hub_rmat24.ncu-rep --page details --csv > gpurun_out/r2_hub_rmat24_details.csv 2>/dev/null
ncu -i /tmp/r2_hub_rmat24.ncu-rep --page raw --csv > gpurun_out/r2_hub_rmat24_raw.csv 2>/dev/null
du -sh gpurun_out
