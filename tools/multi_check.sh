#!/bin/bash
# Multi-GPU check on a box with N GPUs: the multi-GPU parity tests, then bench.py (reference-free) at 1..N GPUs.
#   N=8 bash tools/multi_check.sh
set -u
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -6
for n in ${NS:-1 2 4 8}; do
  [ $n -gt $N ] && continue
  for w in ${WL:-C3}; do
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --workload $w --no-cpu-baseline --no-extras ${EXTRA:-} > gpurun_out/r02_scale_${w}_n$n.json 2> gpurun_out/r02_scale_${w}_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --workload $w --no-cpu-baseline ${EXTRA:-} > gpurun_out/r02_scale_${w}_n$n.json 2> gpurun_out/r02_scale_${w}_n$n.err
    fi
    echo "$w n=$n rc=$? $(python -c "import json; d=json.load(open('gpurun_out/r02_scale_${w}_n$n.json')); print(round(d['value']), 'GCUPS', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value']), 'c2weak', d.get('c2_weak_value') and round(d['c2_weak_value']), '|', d.get('parity_check'))" 2>&1 | tail -1)"
    grep -i "error\|PARITY" gpurun_out/r02_scale_${w}_n$n.err | tail -3
  done
done
