#!/bin/bash
# Device ingest on the GPU box: its tests and the CLI / full-size file tests that now run through it, host-vs-device phase
# times on C2 / C3 / C4, the launch list of one device ingest of the C3 file, and the default bench line with its ingest leg.
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ingest.py tests/test_gpu_cli.py tests/test_gpu_fullsize.py -x -q ) 2>&1 | tail -8
python tools/ingest_bench.py C2 C3 C4 > gpurun_out/r02_ingest_bench.jsonl 2> gpurun_out/r02_ingest_bench.err; cat gpurun_out/r02_ingest_bench.jsonl; tail -3 gpurun_out/r02_ingest_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:"newlines|index_records|gather_ends|pick_" \
    --log-file gpurun_out/r02_ingest_launches_C3.csv python tools/ingest_bench.py C3 --reps 1 > gpurun_out/r02_ingest_ncu.log 2>&1
grep -c . gpurun_out/r02_ingest_launches_C3.csv
( time python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_default.json"))
print({k: d[k] for k in ("value", "ms_per_step", "roofline_frac", "ingest_device_ms", "ingest_host_ms", "parity_check")}, d["e2e"]["value"], d["ingest"])
PY
tail -2 gpurun_out/r02_bench_default.err
