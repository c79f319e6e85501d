#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  OVN_K4_PROD=$v timeout 300 python tools/time_stages.py f16_tc 1101 2>&1 | grep heads > gpurun_out/prod_$v.log
done
OVN_DEBUG_SYNC=1 OVN_K4_PROD=0 timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -q -x -k "heads or full_size" > gpurun_out/pytest_p0.log 2>&1
echo done
