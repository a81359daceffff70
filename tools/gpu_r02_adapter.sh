#!/bin/bash
# the adapter classes driven against tests/mock_ceres on the device
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02_adapter*
timeout 100 python -m pytest tests/test_adapter_mock.py -q -k "adapter_classes" > $OUT/r02_adapter_mock_gpu.log 2>&1
echo "rc=$?" >> $OUT/r02_adapter_mock_gpu.log
