#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02i_*
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/r02i_tests.log
for w in ladybug-1723 venice-1778; do
  echo "== $w" >> $OUT/r02i_prof.log
  B200_VERBOSE=1 timeout 300 python tools/profile_kernels.py $w 10 >> $OUT/r02i_prof.log 2>&1
done
export B200BA_LIB=$PWD/ceres_solver_b200/libb200ba_dev.so
echo "== ladybug-1723 B200_NO_CB3 (gather kernel)" >> $OUT/r02i_prof.log
B200_NO_CB3=1 timeout 300 python tools/profile_kernels.py ladybug-1723 10 2>&1 | grep diag >> $OUT/r02i_prof.log
for mode in default pdl soft; do
  unset B200_CG_PDL B200_CG_SOFT_BARRIER
  if [ $mode = pdl ]; then export B200_CG_PDL=1; fi
  if [ $mode = soft ]; then export B200_CG_SOFT_BARRIER=1; fi
  for rep in 1 2; do
    echo "== bench ladybug-1723 cg=$mode rep=$rep" >> $OUT/r02i_ab.log
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>>$OUT/r02i_ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['final_cost'], d['cg_iterations'])" >> $OUT/r02i_ab.log 2>&1
  done
done
unset B200_CG_PDL B200_CG_SOFT_BARRIER B200BA_LIB
