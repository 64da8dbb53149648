#!/bin/bash
# Round 2, GPU call 1: whole GPU test suite, smoke, then the bench lines (default = C3 strong, C2 weak, C1).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r02a_bench_C3_n1.json 2> gpurun_out/r02a_bench_C3_n1.err; echo "bench C3 rc=$? $(cut -c1-200 gpurun_out/r02a_bench_C3_n1.json)"; tail -3 gpurun_out/r02a_bench_C3_n1.err
python bench.py --workload C2 --scaling weak --steps 30 --no-cpu-baseline > gpurun_out/r02a_bench_C2_n1.json 2> gpurun_out/r02a_bench_C2_n1.err; echo "bench C2 rc=$? $(cut -c1-200 gpurun_out/r02a_bench_C2_n1.json)"
python bench.py --workload C1 --steps 400 --no-cpu-baseline > gpurun_out/r02a_bench_C1_n1.json 2> gpurun_out/r02a_bench_C1_n1.err; echo "bench C1 rc=$? $(cut -c1-200 gpurun_out/r02a_bench_C1_n1.json)"
python bench.py --workload C1 --steps 400 --no-cpu-baseline --no-graph --no-extras > gpurun_out/r02a_bench_C1_nograph.json 2> gpurun_out/r02a_bench_C1_nograph.err; echo "bench C1 nograph rc=$? $(cut -c1-200 gpurun_out/r02a_bench_C1_nograph.json)"
