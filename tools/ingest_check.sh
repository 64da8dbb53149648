#!/bin/bash
# Device ingest on the GPU box: its tests, host-vs-device phase times on C2 / C3 / C4 (file copy through the page-locked
# staging pieces and without), the one-shot binary's wall clock with --ingest device, and the launch list of one
# device ingest of a C2-sized file.
set -u
mkdir -p gpurun_out
nproc
( time python -m pytest tests/test_gpu_ingest.py -x -q ) 2>&1 | tail -8
python tools/ingest_bench.py C2 C3 C4 > gpurun_out/r02_ingest_bench.jsonl 2> gpurun_out/r02_ingest_bench.err; cat gpurun_out/r02_ingest_bench.jsonl; tail -3 gpurun_out/r02_ingest_bench.err
python tools/ingest_bench.py C3 C4 --staging 0 > gpurun_out/r02_ingest_bench_nostaging.jsonl 2>> gpurun_out/r02_ingest_bench.err; cat gpurun_out/r02_ingest_bench_nostaging.jsonl
python tools/cli_wall.py --ingest device C2 C3 C4 > gpurun_out/r02_cli_wall_device_ingest.md 2> gpurun_out/cli_wall.err; cat gpurun_out/r02_cli_wall_device_ingest.md; tail -3 gpurun_out/cli_wall.err
python tools/cli_wall.py C2 C3 C4 > gpurun_out/r02_cli_wall_host_ingest.md 2> gpurun_out/cli_wall.err; cat gpurun_out/r02_cli_wall_host_ingest.md; tail -3 gpurun_out/cli_wall.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_ingest_launches_C2.csv python tools/ingest_bench.py C2 --reps 1 > gpurun_out/r02_ingest_ncu.log 2>&1
grep -c . gpurun_out/r02_ingest_launches_C2.csv
