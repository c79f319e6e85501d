#!/bin/bash
# final-ish artefacts: launch list + ncu of delta kernel at full size + 8-GPU-style bench json
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_delta_conv1_tc -s 1 -c 1 -o gpurun_out/prof_delta_full python tools/time_stages.py f16_tc 1101 > gpurun_out/ncu_delta.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo done
