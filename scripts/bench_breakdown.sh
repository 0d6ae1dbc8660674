#!/bin/bash
# one line per (workload, mode): ms/step, per-tier sweep ms, commit-rule ms, apply+activate ms
for spec in "$@"; do
  w=${spec%%:*}; m=${spec##*:}; [ "$m" = "$w" ] && m=clustering
  timeout 200 python bench.py --workload $w --mode $m --no-cpu-baseline --no-e2e 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['config']['workload'], d['config']['mode'], round(d['ms_per_step'],2), [round(x,2) for x in r['all_sweeps']['per_group_ms']], 'commit', round(r['commit_ms'],2), 'apply', round(r['apply_activate_ms'],2), d['gpu_launches'])"
done
