#!/bin/bash
# One `ncu --set full` capture of the four ingest kernels on the C3 input file (1.1 GB FASTQ): count_newlines,
# write_newlines, index_records, gather_ends — after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
python tools/ingest_bench.py C3 --reps 1 > gpurun_out/ncu_ingest_plain.json 2> gpurun_out/ncu_ingest_plain.err || { tail -5 gpurun_out/ncu_ingest_plain.err; exit 1; }
cat gpurun_out/ncu_ingest_plain.json | cut -c1-200
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"count_newlines|write_newlines|index_records|gather_ends" -c 4 \
    -f -o gpurun_out/r02_ingest_kernels python tools/ingest_bench.py C3 --reps 1 > gpurun_out/ncu_ingest.log 2>&1
ls -la gpurun_out/r02_ingest_kernels.ncu-rep
ncu -i gpurun_out/r02_ingest_kernels.ncu-rep --page raw --csv > gpurun_out/r02_ingest_kernels_raw.csv 2>/dev/null
wc -c gpurun_out/r02_ingest_kernels_raw.csv
