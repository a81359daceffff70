#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02c_*
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -150 > $OUT/r02c_parity.log
B200_VERBOSE=1 timeout 300 python tools/gpu_debug.py tiny > $OUT/r02c_debug_tiny.log 2>&1
B200_VERBOSE=1 timeout 300 python tools/gpu_debug.py circle > $OUT/r02c_debug_circle.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_debug.py circle > $OUT/r02c_sanitizer.log 2>&1
timeout 900 python -m pytest tests/test_gpu_orders.py -m gpu -q 2>&1 | tail -80 > $OUT/r02c_orders.log
