#!/bin/bash
# profiling pass: launch list of one bench step + one full-set capture of each head kernel
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 120 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_delta_conv1_tc|k_conv2_sw_tc|k_conv3_resident_tc|k_corr_tc' -s 4 -c 4 -o gpurun_out/prof_heads_v2 python tools/time_stages.py f16_tc 1101 > gpurun_out/ncu_heads.log 2>&1
echo done
