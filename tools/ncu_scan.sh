#!/bin/bash
# One whole scan under ncu: launch list of a short bench run, then `--set full` over every launch of ONE scan (the
# `start` end of the first timed step) — after the same command has exited 0 without ncu.
#   W=C3 SKIP=66 COUNT=11 bash tools/ncu_scan.sh
set -u
W=${W:-C3}; SKIP=${SKIP:-66}; COUNT=${COUNT:-11}
ARGS="--workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/r02_ncu_plain_$W.json 2> gpurun_out/r02_ncu_plain_$W.err || { echo "plain run failed"; exit 1; }
echo "plain: $(cut -c1-160 gpurun_out/r02_ncu_plain_$W.json)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_$W.csv \
  python bench.py $ARGS > gpurun_out/r02_ncu_launches_$W.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_launches_$W.csv)"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:bs_ -s $SKIP -c $COUNT \
  -o gpurun_out/r02_bs_${W}_scan -f python bench.py $ARGS > gpurun_out/r02_ncu_full_$W.log 2>&1
echo "full rc=$? $(ls -la gpurun_out/r02_bs_${W}_scan.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
# the report itself is too big to travel (64 MiB limit for gpurun_out): keep its raw page (all launches) and the
# SASS source page of the longest launch, drop the report
ncu -i gpurun_out/r02_bs_${W}_scan.ncu-rep --page raw --csv > gpurun_out/r02_bs_${W}_scan_raw.csv 2>/dev/null
python - "$W" <<'PY'
import csv, subprocess, sys
w = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/r02_bs_{w}_scan_raw.csv")))
hdr = rows[0]
dur = hdr.index("gpu__time_duration.sum")
body = [r for r in rows[2:] if len(r) == len(hdr)]
longest = max(range(len(body)), key=lambda i: float(body[i][dur].replace(",", "")))
rid = body[longest][hdr.index("ID")]
print("longest launch: id", rid, body[longest][hdr.index("Kernel Name")][:80], body[longest][dur])
out = subprocess.run(["ncu", "-i", f"gpurun_out/r02_bs_{w}_scan.ncu-rep", "--page", "source", "--csv", "--print-source", "sass",
                      "--launch-skip", str(longest), "--launch-count", "1"], capture_output=True, text=True).stdout
open(f"gpurun_out/r02_bs_{w}_scan_top_kernel_sass.csv", "w").write(out)
PY
rm -f gpurun_out/r02_bs_${W}_scan.ncu-rep
ls -la gpurun_out | tail -8
