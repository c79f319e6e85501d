#!/bin/bash
mkdir -p gpurun_out
export OVN_DEBUG_SYNC=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
unset OVN_DEBUG_SYNC
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench exit=$?" >> gpurun_out/bench2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_delta_conv1_tc -s 1 -c 1 -o gpurun_out/prof_delta python tools/time_stages.py f16_tc 296 > gpurun_out/ncu_delta.log 2>&1
echo done
