#!/bin/bash
# the driver's round-end sequence, twice for flakiness: pytest -m gpu -x, smoke
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02chk_*
for rep in 1 2; do
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > $OUT/r02chk_tests_$rep.log
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02chk_smoke.log 2>&1
B200_VERBOSE=1 timeout 200 python tools/profile_kernels.py ladybug-1723-random 10 > $OUT/r02chk_prof_random.log 2>&1
timeout 200 python tools/profile_kernels.py ladybug-1723 10 > $OUT/r02chk_prof_l1723.log 2>&1
