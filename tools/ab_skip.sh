#!/bin/bash
# A/B of the dead-row skipping: checkpoint row (A/B libraries from tools/build_ab.sh) and the planner's alive share.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
run() { # tag, lib ('' = default), workload, alive
  APC_LIB_PATH=${2:+$PWD/approx_counter_b200/csrc/ab/libapc_$2.so} python bench.py --workload $3 --steps ${5:-10} --warmup 3 --no-cpu-baseline \
    --plan-alive-pct $4 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  echo "$1 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/ab_$1.json'));print(round(d['value']),round(d['ms_per_step'],3),round(d['e2e']['value']))" 2>&1 | tail -1)"
}
run c2_m13_a15 "" C2 15
run c2_m13_a30 "" C2 30
run c2_m13_a50 "" C2 50
run c2_m12_a30 m12 C2 30
run c2_m14_a30 m14 C2 30
run c2_m99_a30 m99 C2 30
run c3_m13_a30 "" C3 30 4
run c3_m12_a30 m12 C3 30 4
run c3_m14_a30 m14 C3 30 4
run c4_m13_a30 "" C4 30 3
run c4_m14_a30 m14 C4 30 3
