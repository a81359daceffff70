#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02mh_*
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -30 > $OUT/r02mh_multirank.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/r02mh_l1723_n2.json 2> $OUT/r02mh_bench.err
tail -1 $OUT/r02mh_l1723_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['final_cost'])" > $OUT/r02mh_bench.log 2>&1
