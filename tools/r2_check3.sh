#!/bin/bash
# GPU call: family kernel A/B on C2 and C3 (parity: fuzz + scan tests first, under a timeout)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
for w in ${WL:-C2 C3}; do
  for fam in "" "--no-family"; do
    timeout 300 python bench.py --workload $w --steps 10 --no-cpu-baseline --no-extras $fam > gpurun_out/r02c_${w}${fam}.json 2> gpurun_out/r02c_${w}${fam}.err
    echo "bench $w $fam rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/r02c_${w}${fam}.json')); print(round(d['value']), round(d['ms_per_step'],3), 'frac', round(d['roofline_frac'],3), 'exec/step', d['lop3_executed_per_step'], 'top', round(d['lop3_top_share_of_executed'],3))" 2>&1 | tail -1)"
  done
done
