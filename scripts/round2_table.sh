#!/bin/bash
# Round-2 bench table on one GPU: every BASELINE config + modes. Raw lines -> gpurun_out/r2_bench_table.jsonl
out=gpurun_out/r2_bench_table.jsonl
: > $out
run() { timeout 900 python bench.py --no-cpu-baseline "$@" >> $out 2>> gpurun_out/r2_bench_table.err || echo "{\"failed\": \"$*\"}" >> $out; }
run --steps 5 --warmup 3                                  # rmat22 clustering (config 2), with e2e
run --workload rmat24 --steps 3 --warmup 3                # config 4 on one GPU
run --workload grid512 --steps 3 --warmup 3               # config 3
run --workload road --steps 5 --warmup 3                  # config 5
run --workload rgg24 --steps 5 --warmup 3                 # rgg2d throughput stand-in
run --mode refinement --steps 5 --warmup 3                # rmat22 k=16 refinement
run --workload grid512 --mode refinement --steps 3 --warmup 3
run --workload road --mode refinement --steps 5 --warmup 3
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_table.jsonl"):
    d = json.loads(ln)
    if "failed" in d:
        print("FAILED", d["failed"]); continue
    c, r = d["config"], d["roofline"]
    print(f'{c["workload"]:8s} {c["mode"]:10s} n={c["n"]:>10} m={c["m_directed"]:>10} {d["ms_per_step"]:8.2f} ms  value {d["value"]/1e9:6.1f} G/s  '
          f'e2e {(d["e2e"] or {}).get("value", 0)/1e9:6.1f} G/s  launches/step {d["gpu_launches"]/d["steps"]:6.0f}  '
          f'dom {r["kernel"]} frac {r["frac"]:.3f} all_sweeps {r["all_sweeps"]["achieved"]:.0f} GB/s')
PY
