#!/bin/bash
# quick 1-GPU check: contraction tests + benches, reference arm with its stderr
timeout 300 python -m pytest tests/test_gpu_contraction.py tests/test_cpp_adapter.py -q -m gpu 2>&1 | tail -3
for w in rmat22 grid256; do
  extra="--no-cpu-baseline"; [ "$w" = "rmat22" ] && extra=""
  timeout 300 python bench.py --workload $w --mode contraction $extra 2>/dev/null | grep '^{' | tail -1 | tee -a gpurun_out/r1_contraction_bench.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'ms', round(d['ms_per_step'],2), 'value %.3g' % d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), d.get('cpu_baseline'))"
done
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_arm.out 2> gpurun_out/ref_arm.err; echo "rc=$?"; cut -c1-300 gpurun_out/ref_arm.out; tail -5 gpurun_out/ref_arm.err
