#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02k_*
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/r02k_tests.log
for w in ladybug-1723 venice-1778 ladybug-1723-random; do
  echo "== $w" >> $OUT/r02k_prof.log
  B200_VERBOSE=1 timeout 300 python tools/profile_kernels.py $w 10 >> $OUT/r02k_prof.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jtj_v4|schur_init_v4|cam_blocks_v2|evaluate_v2|schur_mul_v4' -s 7 -c 5 -o $OUT/r02k_kernels_l1723 python tools/profile_kernels.py ladybug-1723 3 > $OUT/r02k_ncu.log 2>&1
