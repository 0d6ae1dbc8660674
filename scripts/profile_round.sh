#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE GPU). Numbers printed by bench.py under ncu are
# not bench values.
set -u
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|commit_|k_[a-z]" -c 1500 --csv \
  --log-file gpurun_out/r1_final_launches_rmat22.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_a.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"sweep_hub_aggregate|sweep_group|sweep_warp_hash" -s 30 -c 3 \
  -o gpurun_out/r1_final_top python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_b.log 2>&1
ls -la gpurun_out | tail -5
