#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02j_*
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/r02j_tests.log
for w in ladybug-1723 venice-1778; do
  echo "== $w" >> $OUT/r02j_prof.log
  timeout 300 python tools/profile_kernels.py $w 10 >> $OUT/r02j_prof.log 2>&1
done
export B200BA_LIB=$PWD/ceres_solver_b200/libb200ba_dev.so
for mb in 40 79; do
  for w in ladybug-1723 venice-1778; do
    echo "== $w L2 persist $mb MB" >> $OUT/r02j_prof.log
    B200_L2_PERSIST_MB=$mb timeout 300 python tools/profile_kernels.py $w 10 2>&1 | grep -E "L2 persist|schur_|jtj|evaluate|back|model" >> $OUT/r02j_prof.log
  done
done
echo "== bench ladybug-1723 L2 persist 79" >> $OUT/r02j_prof.log
B200_L2_PERSIST_MB=79 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['schur_eliminate'])" >> $OUT/r02j_prof.log 2>&1
unset B200BA_LIB
