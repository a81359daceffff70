#!/bin/bash
# round-2 final single-GPU record: full GPU test suite, smoke, bench lines of the four workloads, launch list, block-SpMV sweep
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02z_*
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/r02z_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02z_smoke.log 2>&1
python bench.py > $OUT/r02z_bench_l1723_default.json 2> $OUT/r02z_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/r02z_bench_l1723_20.json 2>> $OUT/r02z_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload ladybug-1723-random > $OUT/r02z_bench_l1723_random_20.json 2>> $OUT/r02z_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload venice-1778 > $OUT/r02z_bench_v1778_20.json 2>> $OUT/r02z_bench.err
python bench.py --steps 5 --warmup 3 --workload c16 > $OUT/r02z_bench_c16.json 2>> $OUT/r02z_bench.err
timeout 300 python tools/profile_kernels.py ladybug-1723 10 > $OUT/r02z_prof_l1723.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/r02z_launches_l1723.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r02z_launchlist.log 2>&1
timeout 600 python bench.py --workload spmv-sweep --sizes 1e5,3e5,1e6,3e6,1e7,3e7 > $OUT/r02z_spmv.json 2> $OUT/r02z_spmv.err
