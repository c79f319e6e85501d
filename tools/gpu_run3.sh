#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench exit=$?" >> gpurun_out/bench1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo done
