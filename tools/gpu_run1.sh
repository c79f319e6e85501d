#!/bin/bash
# GPU round-trip 1: fp32 parity tests + tcgen05 probes + first timings.  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for m in 0 1 2 3 4 5 6 7 8; do
  timeout 60 ./overlapnet_b200/umma_probe $m >> gpurun_out/probe.log 2>&1
  echo "exit=$?" >> gpurun_out/probe.log
done
timeout 1500 python -m pytest tests -m gpu -q -x -k "not f16_tc and not full_size" > gpurun_out/pytest_fp32.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_fp32.log
timeout 600 python tools/time_stages.py > gpurun_out/time_stages.log 2>&1
echo done
