#!/bin/bash
# (1) the adapter classes driven against tests/mock_ceres on the device; (2) compute-sanitizer memcheck over every entry
# point on the three orderings (direct-mode v4 kernels, re-ordered, id-range family)
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02_memcheck* $OUT/r02_adapter*
timeout 150 python -m pytest tests/test_adapter_mock.py -q -k "adapter_classes" > $OUT/r02_adapter_mock_gpu.log 2>&1
echo "rc=$?" >> $OUT/r02_adapter_mock_gpu.log
timeout 200 compute-sanitizer --tool memcheck --log-file $OUT/r02_memcheck.log \
  python -m pytest tests/test_gpu_orders.py tests/test_gpu_parity.py -q -m gpu -k "every_entry_point or jtj_multiply or schur_multiply or schur_jacobi or evaluate" \
  > $OUT/r02_memcheck_pytest.log 2>&1
echo "rc=$?" >> $OUT/r02_memcheck_pytest.log
