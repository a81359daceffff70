#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02f_*
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/r02f_tests.log
for w in ladybug-1723 venice-1778 ladybug-1723-random; do
  echo "== $w" >> $OUT/r02f_prof.log
  B200_VERBOSE=1 timeout 300 python tools/profile_kernels.py $w 10 >> $OUT/r02f_prof.log 2>&1
done
export B200BA_LIB=$PWD/ceres_solver_b200/libb200ba_dev.so
for lim in 400000 700000; do
  echo "== ladybug-1723-random direct limit $lim" >> $OUT/r02f_prof.log
  B200_DIRECT_LIMIT=$lim B200_VERBOSE=1 timeout 300 python tools/profile_kernels.py ladybug-1723-random 10 >> $OUT/r02f_prof.log 2>&1
done
unset B200BA_LIB
python bench.py --steps 20 --warmup 3 > $OUT/r02f_bench_l1723.json 2> $OUT/r02f_bench.err
