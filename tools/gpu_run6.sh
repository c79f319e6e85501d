#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do
  OVN_K4_VARIANT=$v timeout 300 python tools/time_stages.py f16_tc 1101 2>&1 | grep heads > gpurun_out/variant_$v.log
done
export OVN_DEBUG_SYNC=1
for v in 0 2; do
  OVN_K4_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -x -k "heads or full_size" > gpurun_out/pytest_v$v.log 2>&1
done
echo done
