#!/bin/bash
# A/B of unit-shape masks (bit s = shape s of bs_shape) on the given workloads:  MASKS="0xFFFFFFFF 0x1FFFCF" WL="C2 C3 C4" bash tools/ab_shapes.sh
for w in ${WL:-C2 C3 C4}; do for m in ${MASKS:-0xFFFFFFFF}; do echo "$w mask $m: $(python bench.py --workload $w --steps ${STEPS:-8} --no-cpu-baseline --no-extras --shape-mask $m 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']), round(d['ms_per_step'],3), 'frac', round(d['roofline_frac'],3), 'exec', round(d['lop3_executed_per_step']/1e9,1))")"; done; done
