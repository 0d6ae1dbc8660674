#!/bin/bash
# hub-tier probe: parity of the hub paths, per-tier times on rmat22 / rmat24, per-launch times of the hub kernels
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -4
for w in rmat22 rmat24; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['config']['workload'], 'ms/step', round(d['ms_per_step'],2), 'per-tier ms', [round(x,2) for x in r['all_sweeps']['per_group_ms']], d['config']['moved'])"
done
