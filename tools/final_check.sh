#!/bin/bash
# What the driver runs at round end, in one GPU call: the GPU test suite, smoke(), the reference arm and the default
# bench line (both timed with the shell's clock as well).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1; nproc
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err ) 2>&1 | grep real
echo "reference: $(cut -c1-220 gpurun_out/r02_bench_reference.json)"; tail -2 gpurun_out/r02_bench_reference.err
( time python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
echo "default: $(cut -c1-220 gpurun_out/r02_bench_default.json)"; tail -2 gpurun_out/r02_bench_default.err
python bench.py --workload C2 --scaling weak --steps 30 --no-cpu-baseline > gpurun_out/r02_bench_C2_n1.json 2> gpurun_out/r02_bench_C2_n1.err; echo "C2: $(cut -c1-160 gpurun_out/r02_bench_C2_n1.json)"
python bench.py --workload C1 --steps 400 --no-cpu-baseline > gpurun_out/r02_bench_C1_n1.json 2> gpurun_out/r02_bench_C1_n1.err; echo "C1: $(cut -c1-160 gpurun_out/r02_bench_C1_n1.json)"
python bench.py --workload C4 --steps 10 --no-cpu-baseline > gpurun_out/r02_bench_C4_n1.json 2> gpurun_out/r02_bench_C4_n1.err; echo "C4: $(cut -c1-160 gpurun_out/r02_bench_C4_n1.json)"
