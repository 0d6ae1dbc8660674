#!/bin/bash
# tier-6 (global-table) time per step vs the wave budget (8-byte slots per wave; default 2^28 = one wave)
for w in "" 3000000 6000000 12000000; do
  echo "KMP_HUB_WAVE_SLOTS=${w:-default}"
  env ${w:+KMP_HUB_WAVE_SLOTS=$w} timeout 300 python bench.py --workload ${1:-rmat24} --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  ms/step', round(d['ms_per_step'],2), 'per-tier ms', [round(x,2) for x in r['all_sweeps']['per_group_ms']])"
done
