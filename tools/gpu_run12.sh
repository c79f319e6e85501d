#!/bin/bash
# development loop: GPU tests + bench + stage timings
mkdir -p gpurun_out
export OVN_DEBUG_SYNC=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
unset OVN_DEBUG_SYNC
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
timeout 300 python tools/time_stages.py f16_tc 1101 > gpurun_out/time_tc.log 2>&1
echo done
