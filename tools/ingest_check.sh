#!/bin/bash
# Device ingest on the GPU box: its tests, host-vs-device phase times on C2 / C3 / C4, and the launch list (time + DRAM
# bytes per launch) of one device ingest of the C3 file.
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ingest.py -x -q ) 2>&1 | tail -6
python tools/ingest_bench.py C2 C3 C4 > gpurun_out/r02_ingest_bench.jsonl 2> gpurun_out/r02_ingest_bench.err; cat gpurun_out/r02_ingest_bench.jsonl; tail -3 gpurun_out/r02_ingest_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:"newlines|index_records|gather_ends|pick_" \
    --log-file gpurun_out/r02_ingest_launches_C3.csv python tools/ingest_bench.py C3 --reps 1 > gpurun_out/r02_ingest_ncu.log 2>&1
grep -c . gpurun_out/r02_ingest_launches_C3.csv
