#!/bin/bash
# BASELINE.json config 5: -sn 10k -> 10M, -lim 500 -> 50000 (k=16, sl=100), one GPU; one JSON line per point.
mkdir -p gpurun_out; : > gpurun_out/sweep_c5.jsonl
for n in 10000 100000 1000000 10000000; do
  for q in 500 5000 50000; do
    steps=5; [ $((n * q)) -ge 5000000000 ] && steps=2; [ $((n * q)) -ge 100000000000 ] && steps=1
    python bench.py --workload C2 --reads $n --lim $q --steps $steps --warmup 3 --no-cpu-baseline >> gpurun_out/sweep_c5.jsonl 2>> gpurun_out/sweep_c5.err || echo "{\"failed\": [$n, $q]}" >> gpurun_out/sweep_c5.jsonl
    tail -1 gpurun_out/sweep_c5.jsonl | cut -c1-140
  done
done
