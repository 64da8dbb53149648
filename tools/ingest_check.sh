#!/bin/bash
# Device ingest on the GPU box: its tests, the k = 32 scan tests (checkpoint row 15), the host-vs-device phase
# times on C2 / C3, and the launch list of one device ingest of a C2-sized file.
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ingest.py -x -q ) 2>&1 | tail -15
python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py -x -q -k "32 or fuzz" 2>&1 | tail -3
python tools/ingest_bench.py C2 C3 C4 > gpurun_out/r02_ingest_bench.jsonl 2> gpurun_out/r02_ingest_bench.err; cat gpurun_out/r02_ingest_bench.jsonl; tail -3 gpurun_out/r02_ingest_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_ingest_launches_C2.csv python tools/ingest_bench.py C2 --reps 1 > gpurun_out/r02_ingest_ncu.log 2>&1
grep -c . gpurun_out/r02_ingest_launches_C2.csv
