#!/bin/bash
# ROUND-1 script (kernel names of that round; round 2: scripts/profile_round2.sh, scripts/final_round2.sh).
# ncu evidence for profiles/ (run under gpurun on ONE GPU). Numbers printed by bench.py under ncu are
# not bench values. $1 = tag (file prefix), $2 = "full" to add the --set full capture of the top kernels.
set -u
TAG=${1:-r1_final}
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|commit_|k_[a-z]" -c 4000 --csv \
  --log-file gpurun_out/${TAG}_launches_rmat22.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_a.log 2>&1
if [ "${2:-}" = "full" ]; then
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:"sweep_hub_aggregate|sweep_hub_partial|sweep_group|commit_apply_activate" -s 40 -c 4 \
    -o gpurun_out/${TAG}_top python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/prof_b.log 2>&1
fi
ls -la gpurun_out | tail -5
