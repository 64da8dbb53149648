#!/bin/bash
# Last GPU call of a round: the whole GPU suite, smoke() and the default bench line.
# (compute-sanitizer is closed on this pool; the index ranges of the plane walks are argued in DESIGN.md §3.)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? $(cut -c1-140 gpurun_out/bench_default.json)"
