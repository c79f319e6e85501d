#!/bin/bash
# round-end verification + profiling pass
mkdir -p gpurun_out
export OVN_DEBUG_SYNC=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
unset OVN_DEBUG_SYNC
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 120 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_delta_conv1_tc|k_conv2_sw_tc|k_conv3_resident_tc|k_corr_tc' -s 4 -c 4 -o gpurun_out/prof_heads_v3 python tools/time_stages.py f16_tc 1101 > gpurun_out/ncu_heads.log 2>&1
echo done
