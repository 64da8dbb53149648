#!/bin/bash
# Wall-clock profile of the drop-in binary on synthetic C2 / C3 inputs (phase timestamps from -v 1).
set -e
export TIMEFORMAT="wall %R s"
cd "$(dirname "$0")/.."
BIN=approx_counter_b200/csrc/approx_counter
mkdir -p /tmp/apc_cli
python - <<'PY'
import sys, time; sys.path.insert(0, ".")
from approx_counter_b200 import host
t = time.time(); host.synth_write("/tmp/apc_cli/c2.fa", 1002, 100000, 100); print("gen c2", round(time.time() - t, 1), "s")
t = time.time(); host.synth_write("/tmp/apc_cli/c3.fq", 1003, 1000000, 150, fastq=True); print("gen c3", round(time.time() - t, 1), "s")
PY
ls -la /tmp/apc_cli
for rep in 1 2; do
time $BIN -k 16 -sn 100000 -sl 100 -v 2 -lim 2000 -e /tmp/apc_cli/c2_exact -o /tmp/apc_cli/c2_out /tmp/apc_cli/c2.fa | grep -E "ms\]" | tr '\n' ';' | cut -c1-900; echo
done
time $BIN -k 20 -sn 1000000 -sl 150 -v 2 -lim 5000 -e /tmp/apc_cli/c3_exact -o /tmp/apc_cli/c3_out /tmp/apc_cli/c3.fq | grep -E "ms\]" | tr '\n' ';' | cut -c1-1200; echo

head -3 /tmp/apc_cli/c3_out_0.start
