#!/bin/bash
# Round-2 GPU session 1 (one B200): GPU tests with the parity prints, headline bench + extras, the
# reference arm, the ncu launch list and one full-set capture of the two head kernels that changed.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
grep "\[parity\]" gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench1.err; head -c 3000 gpurun_out/r2_bench1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref exit $?"; cat gpurun_out/r2_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_bench_ncu.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_delta_conv1_tc|k_conv2_sw_tc|k_conv3_resident_tc' -s 6 -c 3 \
  -o gpurun_out/r2_prof_heads python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r2_ncu_heads.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -12
