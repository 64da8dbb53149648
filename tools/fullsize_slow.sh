#!/bin/bash
# Full-size, file-level parity of the drop-in binary on BASELINE C3, C4 and one C5 point (1M reads x 5000 k-mers)
# against the oracle pipeline (tests/test_gpu_fullsize.py, slow cases): minutes of CPU time for the oracle's FM
# index, so they do not run in the default `pytest -m gpu`.  Run on a GPU box:
#   gpurun --timeout 3000 -- 'bash tools/fullsize_slow.sh'
# The log goes to gpurun_out/fullsize_slow.log (copied to profiles/ once green).
mkdir -p gpurun_out
APC_RUN_SLOW=1 python -m pytest tests/test_gpu_fullsize.py -m gpu -v -s -x 2>&1 | tee gpurun_out/fullsize_slow.log | tail -25
