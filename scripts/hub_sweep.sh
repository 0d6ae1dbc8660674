#!/bin/bash
# experiment driver: tier-4 table sizing (KMP_HUB_CAP_PCT, slots per 100 labels) x wave budget
# (KMP_HUB_WAVE_SLOTS, 8-byte slots per wave) on one workload; results in profiles/README.md §5
W=${1:-rmat22}
for cfg in "300 268435456" "200 268435456" "400 268435456" "150 268435456" "300 4194304" "300 8388608"; do
  set -- $cfg
  echo -n "pct=$1 wave_slots=$2: "
  KMP_HUB_CAP_PCT=$1 KMP_HUB_WAVE_SLOTS=$2 timeout 200 python bench.py --workload $W --no-cpu-baseline --no-e2e 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],2), [round(x,2) for x in r['all_sweeps']['per_group_ms']], d['gpu_launches'])"
done
