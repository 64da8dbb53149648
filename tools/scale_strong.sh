#!/bin/bash
# BASELINE.json configs 3 and 4: one job split over 2/4/8 GPUs (strong scaling) + multi-GPU parity tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
run() { # workload gpus
  if [ "$2" = 1 ]; then
    python bench.py --workload $1 --scaling strong --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29700 + $2)) \
      bench.py --workload $1 --scaling strong --gpus $2 --steps 3 --warmup 3
  fi > gpurun_out/strong_$1_n$2.json 2> gpurun_out/strong_$1_n$2.err
  echo "$1 N=$2 rc=$? $(cut -c1-150 gpurun_out/strong_$1_n$2.json)"
}
for n in 1 2 4 8; do run C3 $n; done
for n in 1 8; do run C4 $n; done
