#!/bin/bash
# Pin the oracle to the REAL reference: build /root/reference/approx_counter.cpp against SeqAn (>= 2.4.0,
# reference README.md:14) into oracle/_ref/, run it on the golden inputs with -sn >= #reads (then the sampled
# SET is every eligible read and all results are order-independent, reference :844-848) and diff its four
# output files against the oracle pipeline.
#
#   tools/pin_reference.sh [SEQAN_INCLUDE_DIR]
#
# SeqAn is header-only; pass the directory that contains seqan/index.h if it is not on the default include
# path.  Without SeqAn the build is refused with a message and the pin test reports itself as skipped —
# "parity unpinned" stays true until this script has run green once (record its output in profiles/).
set -u
cd "$(dirname "$0")/.."
INC="${1:-}"
mkdir -p oracle/_ref
probe() { echo '#include <seqan/index.h>' | g++ -std=c++14 ${INC:+-I"$INC"} -E -x c++ - >/dev/null 2>&1; }
if [ ! -f /root/reference/approx_counter.cpp ]; then
  echo "pin_reference: /root/reference/approx_counter.cpp not present on this machine"
elif probe; then
  g++ -std=c++14 -fopenmp -O3 -DNDEBUG ${INC:+-I"$INC"} /root/reference/approx_counter.cpp \
      -o oracle/_ref/approx_counter_ref -lrt && echo "pin_reference: built oracle/_ref/approx_counter_ref"
else
  echo "pin_reference: seqan/index.h not found${INC:+ under $INC} — the reference cannot be built here"
fi
python -m pytest tests/test_ref_pin.py -v -rs
