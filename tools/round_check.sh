#!/bin/bash
# One GPU call: full GPU test suite, the bench lines of C2 (default) / C1 / C3 / C4, then (each only after its
# own command has exited 0 without ncu) the launch list and one full ncu capture of the top kernel on C2.
set -u
mkdir -p gpurun_out
[ "${SKIP_TESTS:-0}" = 1 ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$? $(cut -c1-140 gpurun_out/bench_default.json)"
for w in ${WORKLOADS:-C1 C3 C4}; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$? $(cut -c1-140 gpurun_out/bench_$w.json)"
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2_v8.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_v8.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:bs_group_kernel<16, 8, 4' -s 4 -c 1 -o gpurun_out/bs_c2_v8 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v8.log 2>&1
echo "ncu full rc=$?"
