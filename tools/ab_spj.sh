#!/bin/bash
# A/B of the job size of the bit-sliced scan (super-groups of 1024 reads per job) on C2.
mkdir -p gpurun_out
for spj in 0 1 2 3 4 7; do
  python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline --sg-per-job $spj > gpurun_out/ab_spj$spj.json 2> gpurun_out/ab_spj$spj.err
  echo "spj=$spj rc=$? $(python -c "import json;d=json.load(open('gpurun_out/ab_spj$spj.json'));print(round(d['value']),round(d['ms_per_step'],3),round(d['e2e']['value']))" 2>&1 | tail -1)"
done
