#!/bin/bash
# 2-GPU: sharded parity test, bench with and without the peer-memory exchange
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02e_*
NG=${1:-2}
B200_VERBOSE=1 timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -60 > $OUT/r02e_multirank.log
export B200BA_LIB=$PWD/ceres_solver_b200/libb200ba_dev.so
for w in ladybug-1723 venice-1778; do
  for nx in 0 1; do
    if [ $nx = 1 ]; then export B200_NO_PEER_EXCHANGE=1; else unset B200_NO_PEER_EXCHANGE; fi
    echo "== $w gpus=$NG no_peer_exchange=$nx" >> $OUT/r02e_bench.log
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 10 --warmup 3 --workload $w 2>>$OUT/r02e_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cg_iterations'], d['final_cost'], d['kernels'].get('cg_vector'), d['kernels'].get('schur_multiply'), d['kernels'].get('misc'))" >> $OUT/r02e_bench.log 2>&1
  done
done
unset B200_NO_PEER_EXCHANGE
echo "== 1 GPU reference numbers" >> $OUT/r02e_bench.log
for w in ladybug-1723 venice-1778; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline 2>>$OUT/r02e_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cg_iterations'], d['final_cost'])" >> $OUT/r02e_bench.log 2>&1
done
