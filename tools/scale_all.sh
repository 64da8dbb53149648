#!/bin/bash
# 8-GPU box: multi-GPU parity tests, weak scaling of C2 at N = 1, 2, 4, 8, strong scaling of C3 at N = 8.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
bash tools/scale_run.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 \
  bench.py --workload C3 --scaling strong --gpus 8 --steps 5 --warmup 3 > gpurun_out/strong_C3_n8.json 2> gpurun_out/strong_C3_n8.err
echo "C3 strong N=8 rc=$? $(cut -c1-160 gpurun_out/strong_C3_n8.json)"
