#!/bin/bash
# Bench lines for DESIGN.md §7 (one GPU). Output: gpurun_out/bench_table.jsonl
mkdir -p gpurun_out; : > gpurun_out/bench_table.jsonl
for w in rmat22 rmat24 grid256 rgg20; do
  timeout 400 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 >> gpurun_out/bench_table.jsonl
done
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"sweep_hub" -s 6 -c 9 --csv \
  --log-file gpurun_out/r1_final_hub_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/bench_table.jsonl
