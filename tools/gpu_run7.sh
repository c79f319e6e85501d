#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_delta_conv1_tc -s 1 -c 1 -o gpurun_out/prof_delta python tools/time_stages.py f16_tc 296 > gpurun_out/ncu_delta.log 2>&1
echo done
