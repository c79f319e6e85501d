#!/bin/bash
mkdir -p gpurun_out
export OVN_DEBUG_SYNC=1
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_infer.py -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
unset OVN_DEBUG_SYNC
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
echo done
