#!/bin/bash
# BASELINE.json config 5: -sn 10k -> 10M, -lim 500 -> 50000 (k=16, sl=100), one GPU; one JSON line per point.
# usage: tools/sweep_c5.sh [full]   (default: six points; full: all twelve, ~4 GPU-minutes)
mkdir -p gpurun_out; : > gpurun_out/sweep_c5.jsonl
if [ "$1" = full ]; then
  points="10000:500 10000:5000 10000:50000 100000:500 100000:5000 100000:50000 1000000:500 1000000:5000 1000000:50000 10000000:500 10000000:5000 10000000:50000"
else
  points=${SWEEP_POINTS:-"10000:500 100000:5000 1000000:5000 1000000:50000 10000000:500 10000000:5000"}
fi
for pt in $points; do
  n=${pt%%:*}; q=${pt##*:}
  steps=5; [ $((n * q)) -ge 5000000000 ] && steps=2; [ $((n * q)) -ge 100000000000 ] && steps=1
  python bench.py --workload C2 --reads $n --lim $q --steps $steps --warmup 3 --no-cpu-baseline --no-extras --scaling weak >> gpurun_out/sweep_c5.jsonl 2>> gpurun_out/sweep_c5.err || echo "{\"failed\": [$n, $q]}" >> gpurun_out/sweep_c5.jsonl
  tail -1 gpurun_out/sweep_c5.jsonl | cut -c1-140
done
