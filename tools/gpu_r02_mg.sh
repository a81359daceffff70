#!/bin/bash
# multi-GPU: sharded parity test at every world size the box offers, bench at N = 1, 2, 4, 8 (as the driver's scaling run does)
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
NG=${1:-2}
TAG=r02mg${NG}
rm -f $OUT/${TAG}_*
B200_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -40 > $OUT/${TAG}_multirank.log
for w in ladybug-1723 venice-1778; do
  for n in 1 2 4 8; do
    if [ $n -gt $NG ]; then continue; fi
    echo "== $w gpus=$n" >> $OUT/${TAG}_bench.log
    if [ $n = 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --workload $w --no-cpu-baseline > $OUT/${TAG}_${w}_n$n.json 2>>$OUT/${TAG}_bench.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 --workload $w > $OUT/${TAG}_${w}_n$n.json 2>>$OUT/${TAG}_bench.err
    fi
    tail -1 $OUT/${TAG}_${w}_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['final_cost'], sum(d['cg_iterations']))" >> $OUT/${TAG}_bench.log 2>&1
  done
done
