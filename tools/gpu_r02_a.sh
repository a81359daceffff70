#!/bin/bash
# round-2 GPU call A: full GPU test suite, S*x variant A/B (dev build), ncu captures of every hot kernel
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $OUT/r02a_gputests.log
export B200BA_LIB=$PWD/ceres_solver_b200/libb200ba_dev.so
for w in ladybug-1723 venice-1778; do
  for v in 0 1 2; do
    echo "== $w variant $v" >> $OUT/r02a_ab.log
    B200_VARIANT=$v timeout 300 python tools/profile_kernels.py $w 20 2>&1 | grep -E "schur_multiply|jtj" >> $OUT/r02a_ab.log
  done
  for mb in 48 80 110; do
    echo "== $w variant 0 L2 persist $mb MB" >> $OUT/r02a_ab.log
    B200_L2_PERSIST_MB=$mb timeout 300 python tools/profile_kernels.py $w 20 2>&1 | grep -E "schur_multiply|L2 persist" >> $OUT/r02a_ab.log
  done
done
for v in 0 1 2; do
  echo "== bench ladybug-1723 variant $v" >> $OUT/r02a_ab.log
  B200_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['mean_op_ms'], d['e2e']['value'])" >> $OUT/r02a_ab.log
done
echo "== bench ladybug-1723 variant 0 L2 persist 80" >> $OUT/r02a_ab.log
B200_L2_PERSIST_MB=80 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['mean_op_ms'], d['e2e']['value'])" >> $OUT/r02a_ab.log
unset B200BA_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jtj_v4|schur_init_v2|cam_blocks|row_q|evaluate_v2|cg_vector|schur_mul_v4|invert9|backsub|model_cost' -s 11 -c 9 -o $OUT/r02a_kernels_l1723 python tools/profile_kernels.py ladybug-1723 3 > $OUT/r02a_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cg_vector' -s 60 -c 2 -o $OUT/r02a_cgvec_l1723 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02a_ncu2.log 2>&1
python bench.py --steps 20 --warmup 3 > $OUT/r02a_bench_l1723.json 2> $OUT/r02a_bench.err
