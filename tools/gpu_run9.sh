#!/bin/bash
# profiles for the record: launch list of one bench run + ncu full of the delta kernel and the projection kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_delta_conv1_tc -s 1 -c 1 -o gpurun_out/prof_delta_full python tools/time_stages.py f16_tc 1101 > gpurun_out/ncu_delta.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_project -s 4 -c 2 -o gpurun_out/prof_project python tools/time_stages.py fp32 16 > gpurun_out/ncu_project.log 2>&1
echo done
