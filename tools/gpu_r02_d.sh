#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02d_*
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_orders.py -m gpu -q 2>&1 | tail -120 > $OUT/r02d_parity.log
timeout 900 python -m pytest tests/test_gpu_headline.py -m gpu -q 2>&1 | tail -80 > $OUT/r02d_headline.log
for w in ladybug-1723 ladybug-1723-random venice-1778 venice-1778-random; do
  echo "== $w" >> $OUT/r02d_prof.log
  B200_VERBOSE=1 timeout 300 python tools/profile_kernels.py $w 10 >> $OUT/r02d_prof.log 2>&1
done
export B200BA_LIB=$PWD/ceres_solver_b200/libb200ba_dev.so
for w in ladybug-1723 ladybug-1723-random; do
  for coop in 0 1; do
    echo "== bench $w cooperative=$coop" >> $OUT/r02d_ab.log
    if [ $coop = 1 ]; then export B200_CG_COOPERATIVE=1; else unset B200_CG_COOPERATIVE; fi
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>>$OUT/r02d_ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['mean_op_ms'], d['e2e']['value'], d['kernels'].get('cg_vector'), d['jtj_multiply'], d['schur_eliminate'])" >> $OUT/r02d_ab.log 2>&1
  done
done
unset B200_CG_COOPERATIVE
echo "== bench ladybug-1723-random keep caller order" >> $OUT/r02d_ab.log
B200_KEEP_ORDER=1 timeout 300 python bench.py --workload ladybug-1723-random --steps 10 --warmup 3 --no-cpu-baseline 2>>$OUT/r02d_ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['mean_op_ms'], d['e2e']['value'])" >> $OUT/r02d_ab.log 2>&1
unset B200BA_LIB
python bench.py --steps 20 --warmup 3 > $OUT/r02d_bench_l1723.json 2> $OUT/r02d_bench.err
