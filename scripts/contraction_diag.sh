#!/bin/bash
# diagnostic: calibrate the box with the LP bench, then contraction benches and their ncu launch lists
mkdir -p gpurun_out
scripts/bench_breakdown.sh rmat22
for w in rmat22 grid256; do
  timeout 200 python bench.py --workload $w --mode contraction --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'contraction ms', round(d['ms_per_step'],2))"
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_contract|k_tile|k_coarse|k_map|k_flag|CUB_200802" -c 100 --csv \
    --log-file gpurun_out/diag_contraction_$w.csv python bench.py --workload $w --mode contraction --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python - <<PY
import csv, re, collections
rows=list(csv.reader(open("gpurun_out/diag_contraction_$w.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[hi]
ki,vi,ui=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    name=re.sub(r"\(.*","",r[ki]); name=re.sub(r"<.*","",name).split("::")[-1] or r[ki][:60]
    v=float(r[vi].replace(",","")); v = v/1e3 if r[ui]=="ns" else v
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,a in agg.items(): print("   ", k, a[0], round(a[1],1), "us")
PY
done
