#!/bin/bash
# A/B of the persistent PCG kernel against the multi-kernel PCG on the bench workloads (run under gpurun).
summ='import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"],2), "it/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],2), d["config"]["cg_iterations"], "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],3), {k:(v["launches"],round(v["ms"],2)) for k,v in d["kernels"].items() if k in ("schur_multiply","cg_vector","pcg_persistent")})'
for w in "$@"; do
  echo "== $w persistent"; B200_PCG_PERSISTENT=1 timeout 100 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$summ"
  echo "== $w multi-kernel"; timeout 100 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$summ"
done
