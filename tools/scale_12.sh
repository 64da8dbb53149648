#!/bin/bash
# 2-GPU box: N = 1 and N = 2 bench lines of the current build (weak scaling of C2) + the multi-GPU parity tests.
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
echo "N=1 rc=$? $(cut -c1-150 gpurun_out/scale_n1.json)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
  bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
echo "N=2 rc=$? $(cut -c1-150 gpurun_out/scale_n2.json)"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "not 4 and not 8" 2>&1 | tail -2
