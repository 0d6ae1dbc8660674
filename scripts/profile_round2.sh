#!/bin/bash
# Round-2 ncu evidence (one GPU, under gpurun). Numbers printed by bench.py under ncu are not bench values.
#  1. launch list of the default bench command (every kernel of ours, gpu__time_duration)
#  2. --set full captures of the sweep kernels of R-MAT 22 (LP rounds 1..2) and of the hub kernels of R-MAT 24;
#     the pages are exported to CSV on the box (gpurun_out/ is capped at 64 MiB) and only the small report is kept
set -u
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_|commit_|k_[a-z]|reset_" -c 4000 --csv \
  --log-file gpurun_out/r2_launches_rmat22.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sweep_team|sweep_thread" -s 60 -c 24 \
  -o /tmp/r2_top python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_b.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"sweep_hub" -s 6 -c 6 \
  -o /tmp/r2_hub_rmat24 python bench.py --workload rmat24 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_c.log 2>&1
for r in r2_top r2_hub_rmat24; do
  ncu -i /tmp/$r.ncu-rep --page details --csv > gpurun_out/${r}_details.csv 2>/dev/null
  ncu -i /tmp/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ncu -i /tmp/r2_hub_rmat24.ncu-rep --page source --csv -k regex:sweep_hub_aggregate > gpurun_out/r2_hub_aggregate_source.csv 2>/dev/null
cp /tmp/r2_hub_rmat24.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | tail -9
