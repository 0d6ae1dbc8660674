#!/bin/bash
# experiment driver: KMP_TUNE (bit 0 peek-first shared tables, bit 1 load-first hub tables) x KMP_HUB_CAP_PCT
W=${1:-rmat22}
run() { # tune pct mode
  echo -n "tune=$1 pct=$2 mode=$3: "
  KMP_TUNE=$1 KMP_HUB_CAP_PCT=$2 timeout 200 python bench.py --workload $W --mode $3 --no-cpu-baseline 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],2), [round(x,2) for x in r['all_sweeps']['per_group_ms']], d['gpu_launches'], d['config']['moved'][:2])"
}
run 0 300 clustering
run 1 300 clustering
run 2 300 clustering
run 3 300 clustering
run 3 400 clustering
run 3 600 clustering
run 0 300 refinement
run 3 300 refinement
