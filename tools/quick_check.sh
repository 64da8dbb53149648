#!/bin/bash
# Scan parity tests + the bench lines of C2/C3/C4 (short runs, no CPU baseline): the loop for a kernel change.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -2
for w in C2 C3 C4; do
  python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > gpurun_out/quick_$w.json 2> gpurun_out/quick_$w.err
  echo "$w rc=$? $(python -c "import json;d=json.load(open('gpurun_out/quick_$w.json'));print(round(d['value']),round(d['ms_per_step'],3),round(d['e2e']['value']))" 2>&1 | tail -1)"
done
