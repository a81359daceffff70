#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02h_*
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/r02h_tests.log
for w in ladybug-1723 venice-1778 ladybug-1723-random; do
  echo "== $w" >> $OUT/r02h_prof.log
  B200_VERBOSE=1 timeout 300 python tools/profile_kernels.py $w 10 >> $OUT/r02h_prof.log 2>&1
done
timeout 900 python bench.py --workload spmv-sweep --sizes 1e5,1e6,1e7 > $OUT/r02h_spmv.json 2> $OUT/r02h_spmv.err
python bench.py --steps 20 --warmup 3 > $OUT/r02h_bench_l1723.json 2> $OUT/r02h_bench.err
python bench.py --steps 20 --warmup 3 --workload ladybug-1723-random --no-cpu-baseline > $OUT/r02h_bench_l1723_random.json 2>> $OUT/r02h_bench.err
