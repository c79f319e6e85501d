#!/bin/bash
mkdir -p gpurun_out
export OVN_DEBUG_SYNC=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
unset OVN_DEBUG_SYNC
timeout 300 python tools/time_stages.py f16_tc 1101 > gpurun_out/time_tc.log 2>&1
timeout 300 python tools/time_stages.py fp32 256 > gpurun_out/time_fp32.log 2>&1
echo done
