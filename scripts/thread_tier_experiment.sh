#!/bin/bash
# which degree classes the register-sort kernel (sweep_thread<N>) should take: per-tier ms with KMP_THREAD_MAX_DEG = 16 / 32
for wl in ${1:-rgg24 rmat22 rmat24}; do
  for t in 16 32; do
    echo "$wl KMP_THREAD_MAX_DEG=$t"
    KMP_THREAD_MAX_DEG=$t timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  ms/step', round(d['ms_per_step'],2), 'per-tier ms', [round(x,2) for x in r['all_sweeps']['per_group_ms']], 'M edges', [round(e/1e6) for e in r['all_sweeps']['per_group_edges']])"
  done
done
