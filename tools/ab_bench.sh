#!/bin/bash
# A/B of differently compiled libapc builds (tools/build_ab.sh): the bench line of each on the given workloads.
#   VARIANTS="default base unroll2" WL="C2 C3" bash tools/ab_bench.sh
set -u
mkdir -p gpurun_out
for w in ${WL:-C2 C3}; do
  for v in ${VARIANTS:-default}; do
    lib=""; [ "$v" != default ] && lib="approx_counter_b200/csrc/ab/libapc_$v.so"
    APC_LIB_PATH=$lib timeout 300 python bench.py --workload $w --steps ${STEPS:-15} --no-cpu-baseline --no-extras ${EXTRA:-} > gpurun_out/ab_${w}_$v.json 2> gpurun_out/ab_${w}_$v.err
    echo "$w $v rc=$? $(python -c "import json; d=json.load(open('gpurun_out/ab_${w}_$v.json')); print(round(d['value']), 'GCUPS', round(d['ms_per_step'],4), 'ms frac', round(d['roofline_frac'],3), 'e2e', round(d['e2e']['value']))" 2>&1 | tail -1)"
  done
done
